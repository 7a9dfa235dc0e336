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


def test_unbounded_repetition_through_the_cli(tmp_path):
    """`+ * {m,}` patterns (parametric programs: match_kernel + table) file -> file, single-end and paired, several
    batches; the header of a record may grow by the read's own length, which bk_out_bound accounts for."""
    rnd = random.Random(21)
    n = 30000
    heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=10, max_len=90)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"GATTACA"], head_lens=heads, min_len=0, max_len=90)
    (tmp_path / "r1.fq").write_bytes(fq1)
    (tmp_path / "r2.fq").write_bytes(fq2)
    p1, p2 = "^(?P<UMI>[ATGCN]+)atgccat(?P<SB>[ACGT]*)", "(?P<CB>[ACGT]{3,})gattaca"
    r = run_cli("-m", "1", "--quiet", "extract", "-1", str(tmp_path / "r1.fq"), "-o", str(tmp_path / "o.fq"), "-p", p1, "-e", "1")
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.fq").read_bytes() == bo.extract_se(fq1, p1, 1)
    r = run_cli("-m", "1", "--quiet", "extract", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"),
                "-o", str(tmp_path / "o1.fq"), "-O", str(tmp_path / "o2.fq"), "-p", p1, "-P", p2, "-e", "1")
    assert r.returncode == 0, r.stderr
    e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, 1)
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2


def test_malformed_fastq_is_an_error(tmp_path):
    (tmp_path / "bad.fq").write_bytes(b"@a\nACGT\n+\nIII\n")
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "bad.fq"), "-o", str(tmp_path / "o.fq"), "-p",
                "^(?P<UMI>[ATGC]{2})")
    assert r.returncode == 1 and "malformed" in r.stderr


# ---- streaming reader: record boundaries, carry-over, unequal mates, blank tail ----------------------------------

def test_trailing_blank_lines_and_crlf(tmp_path):
    """seq_io accepts newlines after the last record; CR LF line ends go through the device's re-framing."""
    rnd = random.Random(5)
    fq = ragged_fastq(rnd, 3000, plant=[b"GATTACA"], crlf=True)
    (tmp_path / "r.fq").write_bytes(fq + b"\r\n\n\n")
    pat = "(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})"
    r = run_cli("-m", "1", "--quiet", "extract", "-1", str(tmp_path / "r.fq"), "-o", str(tmp_path / "o.fq"), "-p", pat)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.fq").read_bytes() == bo.extract_se(fq, pat, 1)


def test_mates_of_different_record_length_and_count(tmp_path):
    """10x-style: a short R1 against a long R2 (batches are cut by records, not bytes), and R2 holding more records
    than R1: the run ends with the shorter input (zip, run.rs:172-173)."""
    rnd = random.Random(9)
    n = 40000
    r1 = b"".join(b"@r%d 1\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(28)), b"F" * 28) for i in range(n))
    r2 = b"".join(b"@r%d 2\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(150)), b"F" * 150) for i in range(n + 500))
    (tmp_path / "r1.fq").write_bytes(r1)
    (tmp_path / "r2.fq").write_bytes(r2)
    pat = "^(?P<CB>[ATGCN]{16})(?P<UMI>[ATGCN]{12})"
    r = run_cli("-m", "1", "--quiet", "extract", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"),
                "-o", str(tmp_path / "o1.fq"), "-O", str(tmp_path / "o2.fq"), "-p", pat)
    assert r.returncode == 0, r.stderr
    e1, e2 = bo.extract_pe(r1, r2, pat, None, 1)
    assert (tmp_path / "o1.fq").read_bytes() == e1
    assert (tmp_path / "o2.fq").read_bytes() == e2


def test_truncated_input_is_an_error(tmp_path):
    fq = bksyn.generate(bksyn.CONFIGS["C1"], 1, 0, 1000)
    (tmp_path / "t.fq").write_bytes(fq[:-200])
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "t.fq"), "-o", str(tmp_path / "o.fq"), "-p", CONFIG_PATTERNS["C1"][0])
    assert r.returncode == 1 and "malformed" in r.stderr


# ---- compressed I/O (fastq.rs:71-99,114-124,233-249) --------------------------------------

def _bgzf(data, block=30000):
    """BGZF as bgzip writes it: gzip members with the BC subfield, then the empty end block."""
    import struct
    import zlib
    out = bytearray()
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25)
        out += body + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return bytes(out)


def test_gzip_family_inputs(tmp_path):
    """Everything with the gzip magic is read as a multi-member stream (MultiGzDecoder, fastq.rs:115)."""
    import gzip
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS["C3"], 3, 30000)
    p1 = CONFIG_PATTERNS["C3"][0]
    (tmp_path / "r1.fq.gz").write_bytes(gzip.compress(fq1[:len(fq1) // 3], 1) + gzip.compress(fq1[len(fq1) // 3:], 6))
    (tmp_path / "r2.fq.gz").write_bytes(_bgzf(fq2))
    r = run_cli("-m", "2", "--quiet", "extract", "-1", str(tmp_path / "r1.fq.gz"), "-2", str(tmp_path / "r2.fq.gz"),
                "-o", str(tmp_path / "o1.fq"), "-O", str(tmp_path / "o2.fq"), "-p", p1)
    assert r.returncode == 0, r.stderr
    e1, e2 = bo.extract_pe(fq1, fq2, p1, None, 1)
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
    (tmp_path / "bad.gz").write_bytes(gzip.compress(fq1)[:-1000])
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "bad.gz"), "-o", str(tmp_path / "o3.fq"), "-p", p1)
    assert r.returncode == 1 and "gzip" in r.stderr


@pytest.mark.parametrize("flag", ["--gz", "--bgz", "--mgz"])
def test_compressed_outputs(tmp_path, flag):
    """--gz: gzip members; --bgz selects Mgzip and --mgz selects Bgzf (the reference's swap, fastq.rs:71-79).  The
    decompressed text is what is compared: the reference's own bytes depend on its deflate implementation."""
    import gzip
    import struct
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS["C3"], 4, 20000)
    p1 = CONFIG_PATTERNS["C3"][0]
    (tmp_path / "r1.fq").write_bytes(fq1)
    (tmp_path / "r2.fq").write_bytes(fq2)
    r = run_cli("-m", "1", "--quiet", "extract", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"),
                "-o", str(tmp_path / "o1.gz"), "-O", str(tmp_path / "o2.gz"), "-p", p1, flag)
    assert r.returncode == 0, r.stderr
    e1, e2 = bo.extract_pe(fq1, fq2, p1, None, 1)
    z1, z2 = (tmp_path / "o1.gz").read_bytes(), (tmp_path / "o2.gz").read_bytes()
    assert gzip.decompress(z1) == e1 and gzip.decompress(z2) == e2
    if flag == "--mgz":   # BGZF: BC subfield with the block size, 28-byte empty block at the end
        assert z1[3] == 4 and z1[12:14] == b"BC" and z1[-28:-16] == b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0"
        off, blocks = 0, 0
        while off < len(z1):
            off += struct.unpack_from("<H", z1, off + 16)[0] + 1
            blocks += 1
        assert off == len(z1) and blocks > 2
    if flag == "--bgz":   # mgzip: IG subfield with the member size
        assert z1[3] == 4 and z1[12:14] == b"IG"
        off = 0
        while off < len(z1):
            off += struct.unpack_from("<I", z1, off + 16)[0]
        assert off == len(z1)
    # ... and our own reader takes them back
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "o1.gz"), "-o", str(tmp_path / "again.fq"), "-p",
                "^(?P<UMI>[ATGCN]{5})", "-s")
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "again.fq").read_bytes() == bo.extract_se(e1, "^(?P<UMI>[ATGCN]{5})", 1, skip_trimming=True)


def test_output_to_a_pipe(tmp_path):
    fq = bksyn.generate(bksyn.CONFIGS["C1"], 1, 0, 30000)
    (tmp_path / "r1.fq").write_bytes(fq)
    pat = CONFIG_PATTERNS["C1"][0]
    r = subprocess.run("%s -m 1 --quiet -f extract -1 %s -o /dev/stdout -p '%s' | cat > %s" %
                       (CLI, tmp_path / "r1.fq", pat, tmp_path / "o.fq"), shell=True, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.fq").read_bytes() == bo.extract_se(fq, pat, 1)


def test_lz4_in_and_out(tmp_path):
    """--lz4 writes one LZ4 frame (lz4::EncoderBuilder, fastq.rs:244); inputs with the LZ4 magic are decoded
    (fastq.rs:91,116).  Checked against pyarrow's LZ4 frame codec."""
    pa = pytest.importorskip("pyarrow")
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS["C3"], 6, 25000)
    p1 = CONFIG_PATTERNS["C3"][0]
    (tmp_path / "r1.fq.lz4").write_bytes(pa.compress(fq1, codec="lz4", asbytes=True))
    (tmp_path / "r2.fq").write_bytes(fq2)
    r = run_cli("-m", "2", "--quiet", "extract", "-1", str(tmp_path / "r1.fq.lz4"), "-2", str(tmp_path / "r2.fq"),
                "-o", str(tmp_path / "o1.lz4"), "-O", str(tmp_path / "o2.lz4"), "-p", p1, "--lz4")
    assert r.returncode == 0, r.stderr
    e1, e2 = bo.extract_pe(fq1, fq2, p1, None, 1)
    z1, z2 = (tmp_path / "o1.lz4").read_bytes(), (tmp_path / "o2.lz4").read_bytes()
    assert z1[:4] == b"\x04\x22\x4d\x18" and z1[-4:] == b"\0\0\0\0"
    assert pa.decompress(z1, decompressed_size=len(e1), codec="lz4", asbytes=True) == e1
    assert pa.decompress(z2, decompressed_size=len(e2), codec="lz4", asbytes=True) == e2
    r = run_cli("--quiet", "extract", "-1", str(tmp_path / "o2.lz4"), "-o", str(tmp_path / "again.fq"), "-p",
                "^(?P<UMI>[ATGCN]{5})")
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "again.fq").read_bytes() == bo.extract_se(e2, "^(?P<UMI>[ATGCN]{5})", 1)


def test_input_from_a_pipe_and_empty_input(tmp_path):
    fq = bksyn.generate(bksyn.CONFIGS["C1"], 1, 0, 30000)
    (tmp_path / "r1.fq").write_bytes(fq)
    pat = CONFIG_PATTERNS["C1"][0]
    r = subprocess.run("cat %s | %s -m 1 --quiet extract -1 /dev/stdin -o %s -p '%s'" %
                       (tmp_path / "r1.fq", CLI, tmp_path / "o.fq", pat), shell=True, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.fq").read_bytes() == bo.extract_se(fq, pat, 1)
    (tmp_path / "empty.fq").write_bytes(b"")
    for flag, want in (([], b""), (["--gz"], None)):
        out = tmp_path / ("e" + "".join(flag))
        r = run_cli("--quiet", "extract", "-1", str(tmp_path / "empty.fq"), "-o", str(out), "-p", pat, *flag)
        assert r.returncode == 0, r.stderr
        if want is None:
            import gzip
            assert gzip.decompress(out.read_bytes()) == b""       # a gzip file holds at least one member
        else:
            assert out.read_bytes() == want
