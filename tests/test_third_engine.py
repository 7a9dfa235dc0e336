"""Oracle == oracle: the Python restatement's regex engine (`re`, backtracking, leftmost-first) against two
engines that share no code with it -- the `regex` PyPI module (its own matcher) and the C restatement's VM
(oracle/cpu_extract.c, which executes the expanded alternation literally) -- on the expanded patterns
`get_pattern_with_errors` produces (pattern.rs:93-109) and on random reads.  The reference itself cannot run here
(Rust, no toolchain); three engines agreeing on match / no match, match span and capture spans is the strongest
pin available for rows M, B of SURVEY.md section 8a."""
import random
import re
import zlib

import numpy as np
import pytest

regex = pytest.importorskip("regex")

from oracle import barkit_oracle as bo
from oracle import cport

PATTERNS = [
    ("^(?P<UMI>[ATGCN]{12})", [0, 1]),
    ("^(?P<CB>[ATGCN]{16})atgccat", [0, 1, 2, 9]),
    ("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", [0, 1, 2]),
    ("^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", [0, 1, 2]),
    ("(?P<CB>[ATGCN]{16})atgccat", [0, 1, 2]),
    ("(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})", [0, 1, 2]),
    ("(?P<UMI>[ATGCN]{6})acgt$", [0, 1]),
    ("^n{3}(?P<UMI>[ATGCN]{4})", [0, 1]),
    ("^(atg){2}(?P<UMI>[ACGT]{3})TT", [0, 1]),
    ("(?P<UMI>[A-C]{3}[G-T])acg", [1]),
    ("^(?P<UMI>[ATGCN]{8,10})atgccat", [1]),
    ("^(?P<UMI>[ATGCN]{8,10}?)atgccat", [1]),
    ("(?P<CB>[ACGT]{3,5})gattaca(?P<UMI>[ATGCN]{2,4})", [1]),
    ("^N?(?P<UMI>[ACGT]{6})", [0]),
    # unbounded repetition (the CUDA path's parametric programs, tests/test_extract_gpu.py::test_unbounded_*)
    ("^(?P<UMI>[ATGC]+)", [0]),
    ("^(?P<UMI>[ATGCN]{3,})gattaca", [0, 1]),
    ("(?P<UMI>[ACGT]*)atgccat(?P<CB>[ACGT]{2,})", [1]),
    ("(?P<UMI>[ACGT]*?)atgccat(?P<CB>[ACGT]{2,})", [1]),
    ("(?P<UMI>[ACGT]{1,40})gattaca", [1, 2]),
    ("^N*(?P<CB>[ACGT]{4,})T+?(?P<UMI>[ACGT]{2})$", [0]),
    ("(?P<SB>A+)(?P<UMI>.*)$", [0]),
    # user-written alternation and repeated groups (unfolded into variants by the CUDA path's compiler)
    ("(?P<CB>[ACGT]{6})(atgccat|gattaca)", [0, 1]),
    ("^(?P<UMI>[ACGT]{4})(AC|G|TTT)(?P<CB>[ACGTN]{3})", [0]),
    ("(?P<UMI>[ACGT]{4})(AC){1,3}G", [0]),
    ("^(?P<UMI>[ACGT]{3})(A|CG){1,2}?T", [0]),
    ("(?P<UMI>[ACGT]{4})(A?C){3}", [0]),
    ("(?P<SB>[ACGT]{2})(acgt|(TT|GG){1,2})(?P<UMI>[ACGTN]{4})", [1]),
    # the general form (the CUDA path's backtracking program, tests/test_extract_gpu.py::test_general_form_*)
    ("^(?P<UMI>[ATGCN]{4})+", [0]), ("^(?P<UMI>[ATGCN]+)(A*C){2}", [0]), ("^((?P<UMI>[ATGCN]{4})A){1,2}", [0]),
    ("^(?P<UMI>A)|(?P<CB>C)", [0]), ("^(A|(?P<UMI>C))T", [0]), ("(?P<UMI>[ACGT]{4})(AC)+G", [0]),
    ("(?P<CB>[ACGT]{3,})(atgc)+(?P<UMI>[ACGT]{2})", [0, 1]), ("(?P<UMI>A+?)(C|GT)*?T$", [0]),
    ("^(?P<CB>[ACGT]{2})((?P<UMI>A[CG])|T)+", [0]), ("^(?P<UMI>(AC|A)+)C", [0]), ("(?P<UMI>(A|AC)(C|CA)*)T", [0]),
    ("(?P<UMI>[ACGT]{3})(acgt){2,}|^(?P<CB>N+)", [1]),
]

_PIECES = ["^", "atgc", "gattaca", "ATGC", "N", "[ATGCN]", "[ACGT]", "{3}", "{2}", "(?P<UMI>[ATGCN]{4})", "(?P<CB>[ACGT]{3})",
           ".", "(", ")", "acg", "n"]


def _random_reads(rnd, lits, n=300):
    reads = []
    for _ in range(n):
        L = rnd.choice([0, 3, 8, 17, 23, 35, 36, 60, 90])
        seq = bytearray(rnd.choice(b"ACGTACGTACGTN") for _ in range(L))
        if lits and L > 30 and rnd.random() < 0.7:
            lit = bytearray(rnd.choice(lits).upper().encode())
            for _m in range(rnd.choice([0, 0, 1, 1, 2, 3])):
                lit[rnd.randrange(len(lit))] = rnd.choice(b"ACGTN")
            p = rnd.randrange(0, L - len(lit))
            seq[p:p + len(lit)] = lit
        reads.append(bytes(seq))
    return reads


def _spans(m, names):
    if m is None:
        return None
    return (m.span(0),) + tuple(m.span(nm) for nm in names)


@pytest.mark.parametrize("pattern,ks", PATTERNS)
def test_re_and_regex_modules_agree_on_expanded_patterns(pattern, ks):
    rnd = random.Random(zlib.crc32(pattern.encode()))
    lits = re.findall(r"[atgcn]{3,}", pattern)
    for k in ks:
        rx = bo.BarcodeRegex(pattern, k)
        expanded = bo.BarcodePattern(pattern, k).get_pattern_with_errors().replace("(?<", "(?P<")
        third = regex.compile(expanded.encode(), regex.V0)
        names = rx.get_barcode_types()
        for seq in _random_reads(rnd, lits):
            assert _spans(rx.regex.search(seq), names) == _spans(third.search(seq), names), (pattern, k, seq)


def test_random_pattern_soup_three_engines():
    """Random patterns built from supported pieces: whenever `re` accepts the expanded text, `regex` must accept it
    too and find the same spans on random reads; and the C VM must agree on match start for those that carry a
    capture group (its entry point is the extract call, which needs one)."""
    rnd = random.Random(20240902)
    checked = 0
    for _ in range(600):
        pat = "".join(rnd.choice(_PIECES) for _ in range(rnd.randint(2, 6)))
        k = rnd.choice([0, 1, 2])
        try:
            expanded = bo.BarcodePattern(pat, k).get_pattern_with_errors().replace("(?<", "(?P<")
            a = re.compile(expanded.encode())
        except (re.error, bo.BarkitError):
            continue
        b = regex.compile(expanded.encode(), regex.V0)
        names = list(a.groupindex)
        reads = _random_reads(rnd, re.findall(r"[atgcn]{3,}", pat), n=40)
        for seq in reads:
            assert _spans(a.search(seq), names) == _spans(b.search(seq), names), (pat, k, seq)
        checked += 1
        if names and set(names) <= {"UMI", "CB", "SB"} and "\\" not in pat:
            fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(reads))
            try:
                r = cport.extract(fq, None, pat, None, k, want_match=True)
            except RuntimeError:
                continue   # the VM refuses what it cannot compile; never a wrong answer
            want = np.array([m.start(0) if (m := a.search(s)) else -1 for s in reads], np.int32)
            assert np.array_equal(r["match1"], want), (pat, k)
    assert checked > 100
