"""The oracle's C restatement (oracle/cpu_extract.c) against the Python oracle, which is the
semantic anchor pinned by the reference's known-answer vectors (tests/test_oracle_kat.py).
CPU-only."""
import random

import numpy as np
import pytest

from oracle import barkit_oracle as bo
from oracle import bksyn, cport
from tests.util import CONFIG_PATTERNS, ragged_fastq


@pytest.mark.parametrize("seq,rc", [   # parse.rs:188-193
    (b"", b""), (b"GGGCCCAAATTT", b"AAATTTGGGCCC"), (b"ATGCN", b"NGCAT"),
    (b"AAP", b"ATT"), (b"CCX", b"AGG"), (b"PPP", b"AAA"),
])
def test_reverse_complement(seq, rc):
    assert cport.reverse_complement(seq) == rc


@pytest.mark.parametrize("expected,pattern,max_error", [   # pattern.rs:263-272, :88-91
    ("^(AA.|A.A|.AA)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 1),
    ("^(...)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 3),
    ("^((...))(?P<UMI>[ATGCN]{3})", "^(aaa)(?P<UMI>[ATGCN]{3})", 4),
    ("^(AA.|A.A|.AA)(?P<UMI>[ATGCN]{3})CCC", "^aaa(?P<UMI>[ATGCN]{3})CCC", 1),
    ("^(ATG.|AT.C|A.GC|.TGC)(?<UMI>[ATGCN]{12})", "^atgc(?<UMI>[ATGCN]{12})", 1),
])
def test_expand(expected, pattern, max_error):
    assert cport.expand(pattern, max_error) == expected


def test_get_captures_doctest():   # pattern.rs:211-224
    r = cport.extract(b"@r\nATGCNNNNNNCCC\n+\nABCDEFGHIJKLM\n", None, "^atgc(?<UMI>[ATGCN]{6})", None, 1,
                      skip_trimming=True)
    assert r["out1"] == b"@r UMI:NNNNNN:EFGHIJ\nATGCNNNNNNCCC\n+\nABCDEFGHIJKLM\n"


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_generator_matches_python(name):
    cfg = bksyn.CONFIGS[name]
    for mate in (1, 2):
        for first, n in ((0, 257), (999_999_990_000, 1500)):
            assert cport.generate(cfg, mate, first, n, bksyn.DEFAULT_SEED, threads=3).tobytes() == \
                bksyn.generate(cfg, mate, first, n)


@pytest.mark.parametrize("name,k,skip,rc", [
    ("C1", 1, False, False), ("C2", 1, False, False), ("C3", 1, False, False), ("C4", 0, False, False),
    ("C4", 1, False, False), ("C4", 2, False, False), ("C3", 1, True, False), ("C3", 0, False, True),
    ("C2", 2, True, True),
])
def test_configs_match_python_oracle(name, k, skip, rc):
    p1, p2, paired = CONFIG_PATTERNS[name]
    cfg = bksyn.CONFIGS[name]
    n = 4000
    fq1, fq2 = bksyn.generate_pair(cfg, 31, n)
    if paired:
        e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc, skip)
        for threads in (1, 5):
            r = cport.extract(fq1, fq2, p1, p2, k, rc, skip, threads=threads)
            assert r["out1"] == e1 and r["out2"] == e2 and r["n_in"] == n
    else:
        exp = bo.extract_se(fq1, p1, k, rc, skip)
        r = cport.extract(fq1, None, p1, None, k, rc, skip, threads=3)
        assert r["out1"] == exp


@pytest.mark.parametrize("seed", range(5))
def test_ragged_match_python_oracle(seed):
    rnd = random.Random(seed)
    pats = [("(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})", 1), ("(?P<UMI>[ATGCN]{6})acgt$", 1),
            ("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", 2),
            ("(?P<CB>.{3})[^N]{2}ttagg(?P<UMI>[A-C]{2})", 2), ("^n{3}(?P<UMI>[ATGCN]{4})", 1)]
    pat, k = pats[seed]
    fq = ragged_fastq(rnd, 2000, plant=[b"GATTACA", b"ATGCCAT", b"ACGT", b"TTAGG"], crlf=(seed == 1),
                      plus_comment=True, final_newline=(seed != 2))
    for skip, rc in ((False, False), (True, True)):
        r = cport.extract(fq, None, pat, None, k, rc, skip, threads=4, want_match=True)
        assert r["out1"] == bo.extract_se(fq, pat, k, rc, skip)
    rx = bo.BarcodeRegex(pat, k)
    r = cport.extract(fq, None, pat, None, k, threads=2, want_match=True)
    exp = [(m.start() if m else -1) for m in (rx.regex.search(x.seq) for x in bo.read_fastq(fq))]
    assert list(r["match1"]) == exp


def test_regex_vm_beyond_the_lowered_grammar():
    """The C checker runs the expanded alternation through a real (small) regex VM; exercise
    constructs the CUDA lowering refuses, to show the VM is not the same shortcut."""
    import re
    rnd = random.Random(3)
    for pat in ["(?P<UMI>A+C)", "(?P<UMI>AC|A)(?P<CB>C?G)", "(?P<UMI>[ACG]{2,4})T", "x?(?:a|b)(?P<UMI>.)",
                "(?P<UMI>A|C|G)(?:T|)N*$", "^(?:AC)*(?P<UMI>[GT]{2})"]:
        rx = re.compile(pat.encode())
        recs, exp = [], []
        for i in range(500):
            s = bytes(rnd.choice(b"ACGTNabx") for _ in range(rnd.randint(0, 12)))
            recs.append(b"@r\n" + s + b"\n+\n" + s + b"\n")
            m = rx.search(s)
            exp.append(m.start() if m else -1)
        r = cport.extract(b"".join(recs), None, pat.upper().replace("X?", "x?").replace("(?:A|B)", "(?:a|b)")
                          if False else pat, None, 0, want_match=True)
        # lowercase runs in these patterns would be expanded by the finder; compare with k=0 expansion
        rx2 = re.compile(bo._rust_regex_to_py(bo.BarcodePattern(pat, 0).get_pattern_with_errors()).encode())
        exp = []
        for rec in bo.read_fastq(b"".join(recs)):
            m = rx2.search(rec.seq)
            exp.append(m.start() if m else -1)
        assert list(r["match1"]) == exp, pat


def test_pretokenised_slicing_equals_walk():
    from barkit_b200 import _lib
    cfg = bksyn.CONFIGS["C3"]
    fq1, fq2 = bksyn.generate_pair(cfg, 0, 3000)
    off1, _, _ = _lib.tokenize(fq1)
    off2, _, _ = _lib.tokenize(fq2)
    p1 = CONFIG_PATTERNS["C3"][0]
    a = cport.extract(fq1, fq2, p1, None, 1, threads=4)
    b = cport.extract(fq1, fq2, p1, None, 1, threads=4, off1=off1, off2=off2)
    assert a["out1"] == b["out1"] and a["out2"] == b["out2"]
