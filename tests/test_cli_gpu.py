"""`barkit extract` end to end on a GPU box: files in, files out, byte-identical to the oracle."""
import os
import random
import subprocess

import pytest

from oracle import barkit_oracle as bo
from oracle import bksyn, cport
from tests.util import CONFIG_PATTERNS, ragged_fastq

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "barkit_b200", "barkit")


def run_cli(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def test_single_end_cli(tmp_path):
    fq = bksyn.generate(bksyn.CONFIGS["C1"], 1, 0, 20000)
    (tmp_path / "r1.fq").write_bytes(fq)
    pat = CONFIG_PATTERNS["C1"][0]
    r = run_cli("extract", "-1", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "o1.fq"), "-p", pat)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[:3] == ["[1/3] Estimating reads count...", "[2/3] Parsing barcode patterns...",
                                         "[3/3] Extracting barcodes from reads..."]   # run.rs:97,113,120
    assert "Done in" in r.stdout.splitlines()[3]
    assert (tmp_path / "o1.fq").read_bytes() == bo.extract_se(fq, pat, 1)


@pytest.mark.parametrize("name,k,extra", [("C2", 1, []), ("C3", 1, ["-s"]), ("C4", 2, ["-r"]), ("C4", 0, [])])
def test_paired_end_cli_many_small_batches(tmp_path, name, k, extra):
    """-m 1 forces ~1 MiB batches: several per GPU, so the ordered writer is exercised."""
    p1, p2, _ = CONFIG_PATTERNS[name]
    n = 30000
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS[name], 7, n)
    (tmp_path / "r1.fq").write_bytes(fq1)
    (tmp_path / "r2.fq").write_bytes(fq2)
    args = ["-m", "1", "--quiet", "-f", "extract", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"),
            "-o", str(tmp_path / "o1.fq"), "-O", str(tmp_path / "o2.fq"), "-e", str(k)] + extra
    if p1:
        args += ["-p", p1]
    if p2:
        args += ["-P", p2]
    r = run_cli(*args)
    assert r.returncode == 0, r.stderr
    assert r.stdout == ""                                         # --quiet (logger.rs:68,92)
    ref = cport.extract(fq1, fq2, p1, p2, k, rc_barcodes="-r" in extra, skip_trimming="-s" in extra, threads=4)
    assert (tmp_path / "o1.fq").read_bytes() == ref["out1"]
    assert (tmp_path / "o2.fq").read_bytes() == ref["out2"]


def test_ragged_input_without_final_newline(tmp_path):
    rnd = random.Random(11)
    fq = ragged_fastq(rnd, 5000, plant=[b"GATTACA"], plus_comment=True, final_newline=False)
    (tmp_path / "r.fq").write_bytes(fq)
    pat = "(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})"
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "r.fq"), "-o", str(tmp_path / "o.fq"), "-p", pat, "-e", "1")
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.fq").read_bytes() == bo.extract_se(fq, pat, 1)


def test_malformed_fastq_is_an_error(tmp_path):
    (tmp_path / "bad.fq").write_bytes(b"@a\nACGT\n+\nIII\n")
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "bad.fq"), "-o", str(tmp_path / "o.fq"), "-p",
                "^(?P<UMI>[ATGC]{2})")
    assert r.returncode == 1 and "malformed" in r.stderr
