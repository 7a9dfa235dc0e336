"""Regenerates the golden fixtures of this directory: small bksyn-v1 inputs for every BASELINE.json config and a
ragged hand-made case, with the outputs of the Python restatement of the reference (oracle/barkit_oracle.py,
which follows pattern.rs / parse.rs / run.rs / fastq.rs line by line and passes the reference's own 22
known-answer vectors).  The reference itself is Rust and cannot run in this image, so these are DERIVED
vectors: they pin the C restatement and the CUDA path to the Python one, byte for byte, and guard all three
against drift.

  python tests/golden/make_golden.py      # rewrites tests/golden/*.fq and manifest.json
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import barkit_oracle as bo   # noqa: E402
from oracle import bksyn                 # noqa: E402
from tests.util import CONFIG_PATTERNS, ragged_fastq   # noqa: E402

N = 96          # three tiles, the last one partial
FIRST = 1000    # record index of the first read (bksyn-v1 is counter based)


def write(name, data):
    with open(os.path.join(HERE, name), "wb") as f:
        f.write(data)


def main():
    manifest = []
    for name, (p1, p2, paired) in CONFIG_PATTERNS.items():
        cfg = bksyn.CONFIGS[name]
        for k in ((0, 1, 2) if name == "C4" else (1,)):
            tag = "%s_k%d" % (name.lower(), k)
            if paired:
                fq1, fq2 = bksyn.generate_pair(cfg, FIRST, N)
                o1, o2 = bo.extract_pe(fq1, fq2, p1, p2, k)
                write(tag + "_in1.fq", fq1); write(tag + "_in2.fq", fq2)
                write(tag + "_out1.fq", o1); write(tag + "_out2.fq", o2)
            else:
                fq1 = bksyn.generate(cfg, 1, FIRST, N)
                write(tag + "_in1.fq", fq1)
                write(tag + "_out1.fq", bo.extract_se(fq1, p1, k))
            manifest.append({"tag": tag, "pattern1": p1, "pattern2": p2, "paired": paired, "max_error": k,
                             "rc_barcodes": False, "skip_trimming": False})
    # ragged reads, CR LF line ends, '+' comments, -r and -s
    rnd = random.Random(4711)
    heads = [rnd.choice([3, 23, 60]) for _ in range(80)]
    fq1 = ragged_fastq(rnd, 80, 1, plant=[b"GTCAGTAC"], head_lens=heads, crlf=True)
    fq2 = ragged_fastq(rnd, 80, 2, plant=[b"ATGCCAT"], head_lens=heads, plus_comment=True)
    p1, p2 = "^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", "(?P<CB>[ATGCN]{16})atgccat"
    for tag, rc, skip in (("ragged_rc", True, False), ("ragged_skip", False, True)):
        o1, o2 = bo.extract_pe(fq1, fq2, p1, p2, 1, rc_barcodes=rc, skip_trimming=skip)
        write(tag + "_in1.fq", fq1); write(tag + "_in2.fq", fq2)
        write(tag + "_out1.fq", o1); write(tag + "_out2.fq", o2)
        manifest.append({"tag": tag, "pattern1": p1, "pattern2": p2, "paired": True, "max_error": 1,
                         "rc_barcodes": rc, "skip_trimming": skip})
    # bounded variable repetition (greedy {m,n} and ?): SURVEY.md section 8f-2
    rnd = random.Random(4712)
    heads = [rnd.choice([3, 23, 60]) for _ in range(80)]
    fq1 = ragged_fastq(rnd, 80, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=10, max_len=60)
    fq2 = ragged_fastq(rnd, 80, 2, plant=[b"GATTACA"], head_lens=heads, min_len=0, max_len=60, plus_comment=True)
    p1 = "^(?P<SB>[ATGCN]{0,3})atgccat(?P<UMI>[ATGCN]{4,6})"
    write("varlen_se_in1.fq", fq1)
    write("varlen_se_out1.fq", bo.extract_se(fq1, p1, 1))
    manifest.append({"tag": "varlen_se", "pattern1": p1, "pattern2": None, "paired": False, "max_error": 1,
                     "rc_barcodes": False, "skip_trimming": False})
    p1, p2 = "^N?(?P<UMI>[ACGT]{6,8})", "(?P<CB>[ACGT]{3,5})gattaca(?P<SB>[ACGT]{0,3})"
    o1, o2 = bo.extract_pe(fq1, fq2, p1, p2, 1)
    write("varlen_pe_in1.fq", fq1); write("varlen_pe_in2.fq", fq2)
    write("varlen_pe_out1.fq", o1); write("varlen_pe_out2.fq", o2)
    manifest.append({"tag": "varlen_pe", "pattern1": p1, "pattern2": p2, "paired": True, "max_error": 1,
                     "rc_barcodes": False, "skip_trimming": False})
    # unbounded repetition (* + {m,}): SURVEY.md section 8f-2, the CUDA path's parametric programs
    rnd = random.Random(4713)
    heads = [rnd.choice([3, 23, 60]) for _ in range(80)]
    fq1 = ragged_fastq(rnd, 80, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=10, max_len=60)
    fq2 = ragged_fastq(rnd, 80, 2, plant=[b"GATTACA"], head_lens=heads, min_len=0, max_len=60, plus_comment=True)
    p1 = "^(?P<UMI>[ATGCN]{3,})atgccat(?P<CB>[ACGT]*)"
    write("unbounded_se_in1.fq", fq1)
    write("unbounded_se_out1.fq", bo.extract_se(fq1, p1, 1))
    manifest.append({"tag": "unbounded_se", "pattern1": p1, "pattern2": None, "paired": False, "max_error": 1,
                     "rc_barcodes": False, "skip_trimming": False})
    p1, p2 = "^N*(?P<UMI>[ACGT]+)", "(?P<CB>[ACGT]{2,})gattaca(?P<SB>[ACGT]{0,30})"
    o1, o2 = bo.extract_pe(fq1, fq2, p1, p2, 1)
    write("unbounded_pe_in1.fq", fq1); write("unbounded_pe_in2.fq", fq2)
    write("unbounded_pe_out1.fq", o1); write("unbounded_pe_out2.fq", o2)
    manifest.append({"tag": "unbounded_pe", "pattern1": p1, "pattern2": p2, "paired": True, "max_error": 1,
                     "rc_barcodes": False, "skip_trimming": False})
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote %d cases" % len(manifest))


if __name__ == "__main__":
    main()
