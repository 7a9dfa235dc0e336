"""Parity of the CUDA extract path (through the C ABI) with the oracle.  Needs a B200.

Integer / byte work: the bar is bit-exact output (BASELINE.json north_star).
"""
import ctypes as C
import random
import zlib

import numpy as np
import pytest

from barkit_b200 import _lib
from oracle import barkit_oracle as bo
from oracle import bksyn
from tests.util import CONFIG_PATTERNS, lib_syn_config, ragged_fastq

pytestmark = pytest.mark.gpu


def make_ctx(p1, p2=None, k=1, paired=False, skip=False, rc=False, extra_flags=0, host_threads=None):
    P1 = _lib.Program(p1, k) if p1 is not None else None
    P2 = _lib.Program(p2, k) if p2 is not None else None
    return _lib.Context(P1, P2, paired=paired, skip_trimming=skip, rc_barcodes=rc, extra_flags=extra_flags,
                        host_threads=host_threads)


def run_se(fq, pattern, k=1, skip=False, rc=False, extra_flags=0):
    ctx = make_ctx(pattern, k=k, skip=skip, rc=rc, extra_flags=extra_flags)
    out, res = ctx.extract_text(fq)
    ctx.close()
    if extra_flags == 0:
        # A pattern that needs a search is matched by match_kernel ahead of the extract launch; the same search inside
        # the extract kernel (BK_FLAG_MATCH_IN_KERNEL) must give the same bytes.  (For an anchored pattern the flag
        # changes nothing.)
        out_k, res_k = run_se(fq, pattern, k, skip, rc, extra_flags=_lib.FLAG_MATCH_IN_KERNEL)
        assert out == out_k, "match_kernel + table and the in-kernel search differ"
        assert (res.n_out, res.n_match1) == (res_k.n_out, res_k.n_match1)
    return out, res


def run_pe(fq1, fq2, p1, p2, k=1, skip=False, rc=False, extra_flags=0):
    ctx = make_ctx(p1, p2, k=k, paired=True, skip=skip, rc=rc, extra_flags=extra_flags)
    o1, o2, res = ctx.extract_text(fq1, fq2)
    ctx.close()
    if extra_flags == 0:   # ... and the search inside the extract kernel (see run_se)
        k1, k2, res_k = run_pe(fq1, fq2, p1, p2, k, skip, rc, extra_flags=_lib.FLAG_MATCH_IN_KERNEL | _lib.FLAG_DEVICE_BOTH_MATES)
        assert (o1, o2) == (k1, k2), "match_kernel + table and the in-kernel search differ"
        assert (res.n_out, res.n_match1, res.n_match2) == (res_k.n_out, res_k.n_match1, res_k.n_match2)
    if extra_flags == 0 and p1 is not None and p2 is not None:   # ... and the two-launch form behind match_kernel
        s1, s2, res_s = run_pe(fq1, fq2, p1, p2, k, skip, rc, extra_flags=_lib.FLAG_SPLIT_TWO_PATTERNS)
        assert (o1, o2) == (s1, s2), "paired table launch and one single-end launch per mate differ"
        assert (res.n_out, res.n_match1, res.n_match2) == (res_s.n_out, res_s.n_match1, res_s.n_match2)
    if (p1 is None) != (p2 is None) and extra_flags == 0:
        # One pattern: by default the pattern-less mate is assembled on the host (bk_submit) and the kernel sees a
        # single-end batch.  Every such test also runs the paired kernel (the pattern-less mate copied on the
        # device, global -> global) and must get the same bytes.
        d1, d2, res_d = run_pe(fq1, fq2, p1, p2, k, skip, rc, extra_flags=_lib.FLAG_DEVICE_BOTH_MATES)
        assert (o1, o2) == (d1, d2), "host-side mate and device-side mate differ"
        assert (res.n_out, res.n_match1, res.n_match2) == (res_d.n_out, res_d.n_match1, res_d.n_match2)
    return o1, o2, res


def _qual(n, start=33):
    return bytes(range(start, start + n))


# ---- reference known-answer vectors on the device --------------------------------------------------

@pytest.mark.parametrize("seq,rc", [   # parse.rs:188-193
    (b"", b""), (b"GGGCCCAAATTT", b"AAATTTGGGCCC"), (b"ATGCN", b"NGCAT"),
    (b"AAP", b"ATT"), (b"CCX", b"AGG"), (b"PPP", b"AAA"),
])
def test_get_reverse_complement(seq, rc):
    ctx = make_ctx("^(?P<UMI>[ATGCN]{4})")
    src = np.frombuffer(seq, dtype=np.uint8).copy() if seq else np.zeros(1, np.uint8)
    dst = np.zeros(max(len(seq), 1), dtype=np.uint8)
    assert _lib.lib.bk_reverse_complement(ctx.handle, src.ctypes.data, len(seq), dst.ctypes.data) == 0
    assert dst[:len(seq)].tobytes() == rc
    ctx.close()


def test_get_captures_doctest():   # pattern.rs:211-224
    out, res = run_se(b"@r\nATGCNNNNNNCCC\n+\n" + _qual(13) + b"\n", "^atgc(?<UMI>[ATGCN]{6})", 1, skip=True)
    assert out == b"@r UMI:NNNNNN:" + _qual(13)[4:10] + b"\nATGCNNNNNNCCC\n+\n" + _qual(13) + b"\n"
    assert res.n_out == 1 and res.n_match1 == 1


# ---- derived worked examples ---------------------------------------------------------

def test_derived_examples():
    pat = "^(?P<CB>[ATGCN]{16})atgccat"
    seq = b"ACGTACGTACGTACGT" + b"ATGACAT" + b"GGGGGTTTTT"
    fq = b"@r1 2:N:0:1\n" + seq + b"\n+\n" + _qual(len(seq)) + b"\n"
    out, _ = run_se(fq, pat, 1)
    assert out == bo.extract_se(fq, pat, 1)
    assert out == (b"@r1 2:N:0:1 CB:ACGTACGTACGTACGT:" + _qual(16) + b"\nGGGGGTTTTT\n+\n"
                   + _qual(10, 33 + 23) + b"\n")
    seq2 = b"ACGTACGTACGTACGT" + b"ATGAAAT" + b"GGGGGTTTTT"
    fq2 = b"@r1\n" + seq2 + b"\n+\n" + _qual(len(seq2)) + b"\n"
    assert run_se(fq2, pat, 1)[0] == b""
    assert run_se(fq2, pat, 2)[0] == bo.extract_se(fq2, pat, 2) != b""
    seq3 = b"ACGTACGTACGTACGT" + b"ATGCAT" + b"GGGGGTTTTTT"     # deletion: Hamming only
    fq3 = b"@r1\n" + seq3 + b"\n+\n" + _qual(len(seq3)) + b"\n"
    assert run_se(fq3, pat, 1)[0] == b""
    # B: unanchored, two barcodes, hole cut from the middle
    patb = "(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})"
    seqb = b"NNACGTGATTACAAAAAAACCCCGATTACATTTTTTGG"
    fqb = b"@id desc\n" + seqb + b"\n+\n" + _qual(len(seqb)) + b"\n"
    assert run_se(fqb, patb, 1)[0] == bo.extract_se(fqb, patb, 1)
    # C: reverse-complement quirk
    fqc = b"@q\nTTAAGGGACGT\n+\nABCDEFGHIJK\n"
    assert run_se(fqc, "^(?P<UMI>[ATGCN]{4})ccc", 0, rc=True)[0] == b"@q UMI:TTAA:ABCD\nACGT\n+\nHIJK\n"


def test_readme_header_format():   # the reference's README.md:37
    out, _ = run_se(b"@SEQ_ID\nATGCATGCATGCGG\n+\n????????????II\n", "^(?P<UMI>[ATGC]{4})(?P<CB>[ATGC]{4})(?P<SB>[ATGC]{4})", 0)
    assert out == b"@SEQ_ID UMI:ATGC:???? CB:ATGC:???? SB:ATGC:????\nGG\n+\nII\n"


def test_pe_policy():   # run.rs:149-160
    fq1 = b"@a 1\nACGTTTTT\n+\nIIIIHHHH\n@b 1\nNNNNNNNN\n+x\n12345678\n@c 1\nNAAAAAAA\n+\nabcdefgh\n"
    fq2 = b"@a 2\nGGGGCCCC\n+\nJJJJKKKK\n@b 2\nGGGGAAAA\n+\n87654321\n@c 2\nTTTTTTTT\n+\nhgfedcba\n"
    o1, o2, res = run_pe(fq1, fq2, "^(?P<UMI>[ACGT]{4})", "^(?P<CB>GGGG)", 1)
    assert o1 == b"@a 1 UMI:ACGT:IIII\nTTTT\n+\nHHHH\n@b 1\nNNNNNNNN\n+\n12345678\n"
    assert o2 == b"@a 2 CB:GGGG:JJJJ\nCCCC\n+\nKKKK\n@b 2 CB:GGGG:8765\nAAAA\n+\n4321\n"
    assert (res.n_in, res.n_out, res.n_match1, res.n_match2) == (3, 2, 1, 2)


# ---- BASELINE configs on bksyn-v1 text ------------------------------------------------------------------

@pytest.mark.parametrize("name,k,skip,rc", [
    ("C1", 1, False, False), ("C2", 1, False, False), ("C3", 1, False, False),
    ("C4", 0, False, False), ("C4", 1, False, False), ("C4", 2, False, False),
    ("C3", 1, True, False), ("C3", 0, False, True), ("C2", 2, True, True), ("C1", 1, False, True),
])
def test_baseline_configs_small(name, k, skip, rc):
    p1, p2, paired = CONFIG_PATTERNS[name]
    cfg = bksyn.CONFIGS[name]
    n = 6000 + 37
    if paired:
        fq1, fq2 = bksyn.generate_pair(cfg, 1000, n)
        o1, o2, res = run_pe(fq1, fq2, p1, p2, k, skip, rc)
        e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc, skip)
        assert o1 == e1
        assert o2 == e2
        assert res.n_out == e1.count(b"\n") // 4
    else:
        fq = bksyn.generate(cfg, 1, 1000, n)
        out, res = run_se(fq, p1, k, skip, rc)
        exp = bo.extract_se(fq, p1, k, rc, skip)
        assert out == exp
        assert res.n_out == exp.count(b"\n") // 4


def test_c3_match_rate_is_about_090():
    p1, p2, _ = CONFIG_PATTERNS["C3"]
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS["C3"], 0, 20000)
    _, _, res = run_pe(fq1, fq2, p1, p2, 1)
    assert 0.88 < res.n_out / res.n_in < 0.92


# ---- ragged / edge inputs ----------------------------------------------------------------------------------

@pytest.mark.parametrize("seed", range(6))
def test_ragged_single_end(seed):
    rnd = random.Random(seed)
    pats = [("^(?P<UMI>[ATGCN]{12})", 1), ("(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})", seed % 3),
            ("(?P<UMI>[ATGCN]{6})acgt$", 1), ("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", 1),
            ("(?P<CB>.{3})[^N]{2}ttagg(?P<UMI>[A-C]{2})", 2), ("^(?P<UMI>[ATGCN]{0})", 1)]
    pat, k = pats[seed % len(pats)]
    fq = ragged_fastq(rnd, 3000 + seed, plant=[b"GATTACA", b"ATGCCAT", b"ACGT", b"TTAGG"],
                      crlf=(seed == 2), plus_comment=(seed % 2 == 1), final_newline=(seed != 3))
    for skip, rc in ((False, False), (True, False), (False, True)):
        out, _ = run_se(fq, pat, k, skip, rc)
        assert out == bo.extract_se(fq, pat, k, rc, skip), (pat, k, skip, rc)


@pytest.mark.parametrize("seed", range(4))
def test_ragged_paired_end(seed):
    rnd = random.Random(100 + seed)
    n = 2500 + seed
    heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"GTCAGTAC"], head_lens=heads, min_len=(0 if seed else 30))
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"ATGCCAT"], head_lens=heads, plus_comment=True)
    combos = [("^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", "(?P<CB>[ATGCN]{16})atgccat"),
              (None, "(?P<CB>[ATGCN]{16})atgccat"), ("^(?P<UMI>[ATGCN]{12})", None),
              ("(?P<UMI>[ACGT]{5})gtcagtac", "^(?P<CB>[ATGCN]{16})(?P<SB>[ACGT]{3})")]
    p1, p2 = combos[seed]
    for k in (0, 1, 2):
        o1, o2, _ = run_pe(fq1, fq2, p1, p2, k, skip=(k == 2), rc=(k == 1))
        e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc_barcodes=(k == 1), skip_trimming=(k == 2))
        assert o1 == e1 and o2 == e2, (p1, p2, k)


@pytest.mark.parametrize("period", [1, 2, 3, 5, 31, 32, 33, 1000])
def test_runs_of_unchanged_records(period):
    """Mate 2 is copied verbatim in runs of consecutive kept pairs (warp-cooperative copy); which pairs
    are kept is decided by mate 1.  Every run length / position inside a tile, tiles with no run at all,
    and record lengths that move the runs' ragged ends over all 16 alignments."""
    rnd = random.Random(period)
    recs1, recs2 = [], []
    for i in range(700):
        hit = (i % period) != (period // 2) if period > 1 else True
        if period == 1000:
            hit = i in (0, 31, 32, 63, 699)
        L = rnd.randint(20, 60)
        seq1 = (b"ACGTAC" if hit else b"TTTTTT") + bytes(rnd.choice(b"ACGT") for _ in range(L))
        seq2 = bytes(rnd.choice(b"ACGTN") for _ in range(rnd.randint(1, 70)))
        h = b"r%d" % i
        recs1.append(b"@" + h + b" 1\n" + seq1 + b"\n+\n" + _qual(len(seq1)) + b"\n")
        recs2.append(b"@" + h + b" 2\n" + seq2 + b"\n+\n" + _qual(len(seq2)) + b"\n")
    fq1, fq2 = b"".join(recs1), b"".join(recs2)
    p1 = "^(?P<UMI>ACGTAC)"
    o1, o2, res = run_pe(fq1, fq2, p1, None, 0)
    e1, e2 = bo.extract_pe(fq1, fq2, p1, None, 0)
    assert o1 == e1 and o2 == e2
    # and with the roles of the mates swapped (pattern on mate 2: five back warps there, three here)
    o2, o1, res = run_pe(fq2, fq1, None, p1, 0)
    assert o1 == e1 and o2 == e2


@pytest.mark.parametrize("mode", ["se", "pe1", "pe2", "pe12"])
def test_header_lengths_around_the_scout_window(mode):
    """extract_kernel5 parses and matches inside a 144-byte window at the record start; headers that push
    the pattern (or the header's own line end) past it take the bytewise path.  Sweep header lengths
    0..220 across that boundary, with CR line ends, '+' comments and unequal seq/qual lines mixed in."""
    rnd = random.Random(7)
    n = 2600
    heads = [i % 221 for i in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=20, max_len=90)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"ATGCCAT"], head_lens=heads, plus_comment=True)
    p_long = "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})(?P<SB>[ACGTN]{20})"   # 55 bytes
    p_short = "^(?P<UMI>[ATGCN]{12})"
    if mode == "se":
        for pat in (p_long, p_short, "^(?P<UMI>[ATGCN]{8})acgt$"):
            out, _ = run_se(fq1, pat, 1)
            assert out == bo.extract_se(fq1, pat, 1, False, False), pat
        return
    p1, p2 = {"pe1": (p_long, None), "pe2": (None, p_short), "pe12": (p_short, p_long)}[mode]
    for k, skip in ((1, False), (2, True)):
        o1, o2, _ = run_pe(fq1, fq2, p1, p2, k, skip=skip)
        e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc_barcodes=False, skip_trimming=skip)
        assert o1 == e1 and o2 == e2, (mode, k)


def test_crlf_paired_anchored():
    """CR line ends on both mates with `^` patterns: every record leaves the writer's framing."""
    rnd = random.Random(11)
    n = 1500
    heads = [rnd.choice([0, 1, 23, 130]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT"], head_lens=heads, crlf=True)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"ATGCCAT"], head_lens=heads, crlf=True, final_newline=False)
    p1, p2 = "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", "^(?P<SB>[ATGCN]{5})"
    o1, o2, _ = run_pe(fq1, fq2, p1, p2, 1)
    e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, 1, rc_barcodes=False, skip_trimming=False)
    assert o1 == e1 and o2 == e2


@pytest.mark.parametrize("which", ["p1", "p2"])
def test_host_side_mate_matches_the_device_path(which):
    """Paired mode with one pattern: the pattern-less mate is assembled on the host from the caller's input
    buffer (bk_submit / bk_wait, several threads above 50 000 records).  Same bytes as with both mates on
    the device (BK_FLAG_DEVICE_BOTH_MATES) and as the oracle; CR line ends, '+' comments and a last record
    without EOL included, so that re-framed records shift the parts of the threaded copy."""
    rnd = random.Random(21)
    n = 61000
    heads = [rnd.choice([5, 23, 40]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=30, max_len=60, plus_comment=True)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"ATGCCAT"], head_lens=heads, min_len=30, max_len=60, plus_comment=True,
                       final_newline=False)
    # sprinkle CR line ends over some records of both mates
    def crlf_some(fq):
        recs = bo.read_fastq(fq)
        out = []
        for i, r in enumerate(recs):
            eol = b"\r\n" if i % 97 == 0 else b"\n"
            out.append(b"@" + r.head + eol + r.seq + eol + b"+" + (r.head if i % 5 == 0 else b"") + eol + r.qual + eol)
        return b"".join(out)
    fq1, fq2 = crlf_some(fq1), crlf_some(fq2)[:-1]
    pat = "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{6})"
    p1, p2 = (pat, None) if which == "p1" else (None, pat)
    o1, o2, res = run_pe(fq1, fq2, p1, p2, 1)
    d1, d2, res_d = run_pe(fq1, fq2, p1, p2, 1, extra_flags=_lib.FLAG_DEVICE_BOTH_MATES)
    assert res_d.h2d_bytes > 1.9 * res.h2d_bytes          # the second mate never crossed the bus
    assert (o1, o2) == (d1, d2)
    assert (res.n_out, res.n_match1, res.n_match2) == (res_d.n_out, res_d.n_match1, res_d.n_match2)
    e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, 1, rc_barcodes=False, skip_trimming=False)
    assert o1 == e1 and o2 == e2


VARLEN_PATTERNS = [
    ("^(?P<UMI>[ATGCN]{8,10})atgccat", 1),            # greedy: the longest barcode the linker allows
    ("^(?P<UMI>[ATGCN]{8,10}?)atgccat", 1),           # lazy: the shortest
    ("(?P<CB>[ACGT]{3,5})gattaca(?P<UMI>[ATGCN]{2,4})", 1),   # unanchored, two ranges
    ("^N?(?P<UMI>[ACGT]{6})", 0),                     # optional byte
    ("(?P<UMI>[ATGCN]{4,6})acgt$", 1),                # anchored at the end
    ("^(?P<SB>[ACGT]{0,2})(?P<UMI>[ATGCN]{3})T{1,2}", 0),     # a capture that may be empty
]


@pytest.mark.parametrize("pat,k", VARLEN_PATTERNS)
def test_variable_repetition_single_end(pat, k):
    """Bounded {m,n} / ? repetition: the pattern runs as its fixed-length variants in
    the regex crate's backtracking order; output must equal the oracle's (Python `re`, same leftmost-first
    semantics), with trimming, -s and -r."""
    rnd = random.Random(zlib.crc32(pat.encode()))
    fq = ragged_fastq(rnd, 3000, plant=[b"GATTACA", b"ATGCCAT", b"ACGT"], crlf=False, plus_comment=True,
                      min_len=0, max_len=80)
    for skip, rc in ((False, False), (True, False), (False, True)):
        out, res = run_se(fq, pat, k, skip, rc)
        assert out == bo.extract_se(fq, pat, k, rc, skip), (pat, k, skip, rc)
        assert 0 < res.n_out < 3000


def test_variable_repetition_paired_end():
    rnd = random.Random(5)
    n = 2500
    heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=10, max_len=70)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"GATTACA"], head_lens=heads, min_len=0, max_len=70)
    combos = [("^(?P<UMI>[ATGCN]{8,10})atgccat", "(?P<CB>[ACGT]{3,5})gattaca"),
              ("^(?P<UMI>[ATGCN]{8,10})atgccat", None), (None, "(?P<CB>[ACGT]{3,5}?)gattaca(?P<SB>[ACGT]{0,3})"),
              ("^(?P<UMI>[ATGCN]{12})", "^(?P<CB>[ACGT]{3,5})gattaca")]
    for p1, p2 in combos:
        for k, skip, rc in ((1, False, False), (2, True, False), (0, False, True)):
            o1, o2, _ = run_pe(fq1, fq2, p1, p2, k, skip=skip, rc=rc)
            e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc_barcodes=rc, skip_trimming=skip)
            assert o1 == e1 and o2 == e2, (p1, p2, k, skip, rc)


ALTERNATION = [
    ("(?P<CB>[ACGT]{6})(atgccat|gattaca)", 1), ("^(?P<UMI>[ACGT]{4})(AC|G|TTT)(?P<CB>[ACGTN]{3})", 0),
    ("(?P<UMI>[ACGT]{4})(AC){1,3}G", 0), ("^(?P<UMI>[ACGT]{3})(A|CG){1,2}?T", 0), ("(?P<UMI>[ACGT]{4})(A?C){3}", 0),
    ("(?P<UMI>[ACGT]{4})([AT]{1,2}G){2}", 0), ("(?P<UMI>AC|G[ACGT]{2})gattaca$", 2), ("^(?P<UMI>[ACGT]{4})((A|C)T){2}(?P<CB>[ACGT]{2,3})", 0),
    ("(?P<SB>[ACGT]{2})(acgt|(TT|GG){1,2})(?P<UMI>[ACGTN]{4})", 1),
]


@pytest.mark.parametrize("pat,k", ALTERNATION)
def test_alternation_and_repeated_groups(pat, k):
    """User-written `|` and repetition of groups: the pattern is unfolded into fixed-length
    variants in the regex crate's decision order; output must equal the oracle's (Python `re`: same backtracking
    priority), with trimming, -s and -r, single-end and as either mate of a pair."""
    rnd = random.Random(zlib.crc32(pat.encode()))
    fq = ragged_fastq(rnd, 3000, plant=[b"GATTACA", b"ATGCCAT", b"ACGT", b"ACACAC"], plus_comment=True, min_len=0, max_len=80)
    for skip, rc in ((False, False), (True, False), (False, True)):
        out, res = run_se(fq, pat, k, skip, rc)
        assert out == bo.extract_se(fq, pat, k, rc, skip), (pat, k, skip, rc)
        assert 0 < res.n_out <= 3000
    heads = [rnd.choice([3, 23, 60]) for _ in range(1500)]
    fq1 = ragged_fastq(rnd, 1500, 1, plant=[b"GATTACA", b"ACGT"], head_lens=heads, min_len=5, max_len=70)
    fq2 = ragged_fastq(rnd, 1500, 2, plant=[b"ATGCCAT", b"ACACAC"], head_lens=heads, min_len=0, max_len=70)
    for p1, p2 in ((pat, "^(?P<CB>[ACGT]{5})"), ("(?P<SB>[ACGT]{3})gattaca", pat), (None, pat)):
        o1, o2, _ = run_pe(fq1, fq2, p1, p2, k)
        assert (o1, o2) == bo.extract_pe(fq1, fq2, p1, p2, k), (p1, p2, k)


def _random_pattern(rnd):
    """A random pattern over the lowered grammar: classes and literals with fixed, bounded and unbounded repetition
    (greedy or lazy), lowercase fuzzy runs, alternation groups, repeated groups, one to three named captures."""
    names = ["UMI", "CB", "SB"]
    rnd.shuffle(names)
    ncap = rnd.randint(1, 3)
    cls = lambda: rnd.choice(["[ACGT]", "[ACGTN]", "[AC]", "[GT]", ".", "[^N]"])
    quant = lambda: rnd.choice(["", "", "{2}", "{3}", "{1,3}", "{2,4}?", "{0,2}", "?", "+", "*", "{2,}", "+?", "{1,30}"])
    def atom():
        r = rnd.random()
        if r < 0.35:
            return cls() + quant()
        if r < 0.55:
            return rnd.choice(["A", "C", "G", "T", "AC", "GT", "TTA"]) + (rnd.choice(["", "", "{2}", "{1,2}"]) if rnd.random() < 0.3 else "")
        if r < 0.75:
            return rnd.choice(["acgt", "gattaca", "tt", "atgc", "gg"])
        if r < 0.9:
            return "(" + "|".join(rnd.choice(["A", "CG", "TT", "acg", "G[AC]", "[GT]{2}"]) for _ in range(rnd.randint(2, 3))) + ")" + rnd.choice(["", "", "{2}", "{1,2}", "?"])
        return "(" + rnd.choice(["AC", "G[ACGT]", "T?A", "[AC]{1,2}G"]) + ")" + rnd.choice(["{2}", "{1,2}", "{0,2}?", "{1,3}"])
    parts = []
    for i in range(ncap):
        for _ in range(rnd.randint(0, 2)):
            parts.append(atom())
        inner = cls() + rnd.choice(["{2}", "{4}", "{3,5}", "{2,}", "+", "*", "{1,3}?", "{6}"])
        if rnd.random() < 0.2:
            inner += rnd.choice(["A", "[CG]", "(A|T)"])
        parts.append("(?P<%s>%s)" % (names[i], inner))
    for _ in range(rnd.randint(0, 2)):
        parts.append(atom())
    return ("^" if rnd.random() < 0.5 else "") + "".join(parts) + ("$" if rnd.random() < 0.15 else "")


def test_random_patterns_against_the_oracle():
    """Pattern fuzzing on the device: whatever the compiler accepts must give the oracle's bytes (Python `re` on the
    reference's expanded text), whatever it does not lower must be refused, never approximated."""
    rnd = random.Random(20241020)
    fq = ragged_fastq(rnd, 700, plant=[b"GATTACA", b"ACGT", b"ATGC", b"ACACGG", b"TTATTA"], min_len=0, max_len=48)
    accepted = refused = 0
    for it in range(260):
        pat = _random_pattern(rnd)
        k = rnd.choice([0, 1, 1, 2])
        skip, rc = rnd.random() < 0.25, rnd.random() < 0.25
        try:
            exp = bo.extract_se(fq, pat, k, rc, skip)
        except Exception:
            continue      # (not a valid pattern for the oracle's engine either)
        try:
            ctx = make_ctx(pat, k=k, skip=skip, rc=rc)
        except _lib.BarkitError as e:
            assert e.status in (_lib.BK_E_UNSUPPORTED, _lib.BK_E_PATTERN), (pat, str(e))
            refused += 1
            continue
        try:
            out, _ = ctx.extract_text(fq)
        finally:
            ctx.close()
        assert out == exp, (pat, k, skip, rc)
        accepted += 1
    assert accepted >= 100, (accepted, refused)


UNBOUNDED = [
    ("^(?P<UMI>[ATGC]+)", 0), ("^(?P<UMI>[ATGCN]{3,})gattaca", 1), ("(?P<UMI>[ACGT]*?)atgccat(?P<CB>[ACGT]{2,})", 1),
    ("^N*(?P<CB>[ACGT]{4,})T+?(?P<UMI>[ACGT]{2})$", 0), ("(?P<UMI>[ACGT]{1,40})gattaca", 2), ("(?P<SB>A+)(?P<UMI>.*)$", 0),
    ("^(?P<UMI>.{2,}?)acgt(?P<CB>[ATGCN]*)", 1), ("(?P<CB>[ATGCN]{5,})atgccat(?P<UMI>[ATGCN]{0,30})", 1),
    ("^(?P<UMI>[ATGC]{2,6}?)G*C{1,}(?P<CB>[ACGTN]+)", 0),
]


@pytest.mark.parametrize("pat,k", UNBOUNDED)
def test_unbounded_repetition_single_end(pat, k):
    """`* + {m,}` and widely bounded `{m,n}` on a class, `.` or a literal byte: the pattern
    is one parametric program whose insertion points match_kernel fills by backtracking in the regex crate's order
    (greedy: longest first, lazy: shortest first, the first quantifier outermost), and the extract launch rewrites
    every record with ITS capture offsets and lengths.  Output must equal the oracle's (Python `re`: the same
    backtracking semantics), with trimming, -s and -r."""
    assert _lib.Program(pat, k).nvariants() == 1
    rnd = random.Random(zlib.crc32(pat.encode()))
    fq = ragged_fastq(rnd, 3000, plant=[b"GATTACA", b"ATGCCAT", b"ACGT"], crlf=False, plus_comment=True,
                      min_len=0, max_len=80)
    for skip, rc in ((False, False), (True, False), (False, True)):
        out, res = run_se(fq, pat, k, skip, rc)
        assert out == bo.extract_se(fq, pat, k, rc, skip), (pat, k, skip, rc)
        assert 0 < res.n_out <= 3000


def test_unbounded_repetition_paired_end():
    rnd = random.Random(11)
    n = 2500
    heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT"], head_lens=heads, min_len=10, max_len=70)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"GATTACA"], head_lens=heads, min_len=0, max_len=70, crlf=True)
    combos = [("^(?P<UMI>[ATGCN]+)atgccat", "(?P<CB>[ACGT]{3,})gattaca"),
              ("^(?P<UMI>[ATGCN]{8,})atgccat", None), (None, "(?P<CB>[ACGT]*?)gattaca(?P<SB>[ACGT]+)"),
              ("^(?P<UMI>[ATGCN]{12})", "^(?P<CB>[ACGT]+)gattaca"), ("(?P<UMI>[ATGCN]{4})atgccat", "(?P<CB>N*[ACGT]{2,}?)gattaca")]
    for p1, p2 in combos:
        for k, skip, rc in ((1, False, False), (2, True, False), (0, False, True)):
            o1, o2, _ = run_pe(fq1, fq2, p1, p2, k, skip=skip, rc=rc)
            e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc_barcodes=rc, skip_trimming=skip)
            assert o1 == e1 and o2 == e2, (p1, p2, k, skip, rc)


GENERAL = [
    ("^(?P<UMI>[ATGCN]{4})+", 0), ("^(?P<UMI>[ATGCN]+)(A*C){2}", 0), ("^((?P<UMI>[ATGCN]{4})A){1,2}", 0),
    ("^(?P<UMI>[ATGCN]{4})?", 0), ("^(?P<UMI>A)|(?P<CB>C)", 0), ("^(A|(?P<UMI>C))T", 0), ("^(?P<UMI>[ACGT]+)(A|C)", 0),
    ("(?P<UMI>[ACGT]{4})(AC)+G", 0), ("(?P<CB>[ACGT]{3,})(atgc)+(?P<UMI>[ACGT]{2})", 1), ("(?P<UMI>A+?)(C|GT)*?T$", 0),
    ("^(?P<CB>[ACGT]{2})((?P<UMI>A[CG])|T)+", 0), ("(?P<UMI>[ACGT]{4})(atgccat|gattaca)+", 1), ("^(?P<UMI>(AC|A)+)C", 0),
    ("(?P<UMI>(A|AC)(C|CA)*)T", 0), ("(?P<SB>[ACGT]{2})(gattaca){1,}?(?P<UMI>[ACGT]{3})", 2),
    ("(?P<UMI>[ACGT]{3})(acgt){2,}|^(?P<CB>N+)", 1), ("^(?P<UMI>[ACGT]{4})(A|C|G|T|N){3}", 0),
    ("(?P<CB>[ACGT](?P<UMI>[AC]+)G)(T|$)", 0),
]


@pytest.mark.parametrize("pat,k", GENERAL)
def test_general_form_single_end(pat, k):
    """What the fixed, variant and parametric forms cannot express -- unbounded repetition of
    groups and lowercase runs, alternation at the top level, named groups that may take no part in a match (the read is
    then dropped: parse.rs:127-136), nested named groups -- runs as a backtracking program in match_kernel<VM>
    (match_vm), one lane per read, and the extract launch rewrites every record from the spans it leaves.  Output must
    equal the oracle's (Python `re`: the same priorities), with trimming, -s and -r."""
    assert "vm 0 " in _lib.Program(pat, k).dump()
    rnd = random.Random(zlib.crc32(pat.encode()))
    fq = ragged_fastq(rnd, 3000, plant=[b"GATTACA", b"ATGCCAT", b"ACGT", b"ACACAC", b"ACGTACGT"], plus_comment=True,
                      min_len=0, max_len=80)
    for skip, rc in ((False, False), (True, False), (False, True)):
        out, res = run_se(fq, pat, k, skip, rc)
        assert out == bo.extract_se(fq, pat, k, rc, skip), (pat, k, skip, rc)
        assert res.n_out == out.count(b"\n") // 4       # (a pattern one of whose groups never takes part keeps nothing)


def test_general_form_paired_end():
    rnd = random.Random(12)
    n = 2500
    heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"ATGCCAT", b"ACAC"], head_lens=heads, min_len=10, max_len=70)
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"GATTACA"], head_lens=heads, min_len=0, max_len=70, crlf=True)
    combos = [("^(?P<UMI>[ATGCN]{4})+atgccat", "(?P<CB>[ACGT]{3,})(gattaca)+"), ("(?P<UMI>[ACGT]{4})(AC)+G", None),
              (None, "(?P<CB>[ACGT]*?)(gattaca|ACGT)+(?P<SB>[ACGT]+)"), ("^(?P<UMI>[ATGCN]{12})", "^(?P<CB>[ACGT]+)(gattaca)*T"),
              ("(?P<UMI>[ATGCN]{4})atgccat", "^(?P<CB>A)|(?P<SB>C[GT]+)"), ("^(?P<UMI>[ATGCN]+)atgccat", "(?P<CB>(AC|GT)+)$")]
    for p1, p2 in combos:
        for k, skip, rc in ((1, False, False), (2, True, False), (0, False, True)):
            o1, o2, _ = run_pe(fq1, fq2, p1, p2, k, skip=skip, rc=rc)
            e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc_barcodes=rc, skip_trimming=skip)
            assert o1 == e1 and o2 == e2, (p1, p2, k, skip, rc)


def test_general_form_gives_up_instead_of_guessing():
    """The reference's engine is linear-time; the device matcher is a memoised backtracker whose visited states (marked
    instructions x read length) must fit a fixed number of bits per read.  A read that needs more is refused
    (BK_E_UNSUPPORTED from the batch), never answered differently -- and a pattern that is exponential for a plain
    backtracker is not a problem below that size."""
    pat = "^(?P<UMI>(A|AA)+)(C|G)*(T|N)*$"
    ok = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"A" * (3 + i % 9), b"I" * (3 + i % 9)) for i in range(64))
    ok += b"@x\n%sCT\n+\n%sII\n" % (b"A" * 150, b"I" * 150)     # (one that matches: the ORACLE's engine is exponential on a miss)
    out, _ = run_se(ok, pat, 0)
    assert out == bo.extract_se(ok, pat, 0)
    bad = ok + b"@y\n%sC\n+\n%sI\n" % (b"A" * 3000, b"I" * 3000)
    with pytest.raises(_lib.BarkitError) as e:
        run_se(bad, pat, 0)
    assert e.value.status == _lib.BK_E_UNSUPPORTED and "backtracking" in str(e.value)


def _random_general_pattern(rnd):
    """Random patterns that need the general form: repeated groups without a bound, alternatives at the top level,
    optional and nested named groups, repeated lowercase runs."""
    names = ["UMI", "CB", "SB"]
    rnd.shuffle(names)
    cls = lambda: rnd.choice(["[ACGT]", "[ACGTN]", "[AC]", "[GT]", ".", "[^N]"])
    small = lambda: rnd.choice(["A", "CG", "TT", "acg", "G[AC]", "[GT]{2}", "gatt", "[AC]+T", "C?G"])
    def group():
        body = "|".join(small() for _ in range(rnd.randint(1, 3)))
        return "(" + body + ")" + rnd.choice(["+", "*", "+?", "{2,}", "{1,}?", "{1,3}", ""])
    def cap(i):
        inner = cls() + rnd.choice(["{2}", "{4}", "{3,5}", "{2,}", "+", "*", "{1,3}?"])
        if rnd.random() < 0.3:
            inner += group()
        return "(?P<%s>%s)" % (names[i], inner) + ("?" if rnd.random() < 0.15 else "")
    def branch(first):
        parts = []
        for i in range(first, first + rnd.randint(1, 2)):
            if rnd.random() < 0.6:
                parts.append(group())
            parts.append(cap(i))
        if rnd.random() < 0.7:
            parts.append(group())
        return ("^" if rnd.random() < 0.4 else "") + "".join(parts) + ("$" if rnd.random() < 0.15 else ""), first + 2
    b1, nxt = branch(0)
    if rnd.random() < 0.25:
        return b1 + "|" + "(?P<%s>%s+)" % (names[2], cls())
    return b1


def test_random_general_patterns_against_the_oracle():
    rnd = random.Random(20241021)
    fq = ragged_fastq(rnd, 600, plant=[b"GATTACA", b"ACGT", b"GATT", b"ACGACG", b"TTATTA"], min_len=0, max_len=48)
    accepted = general = gave_up = 0
    for it in range(220):
        pat = _random_general_pattern(rnd)
        k = rnd.choice([0, 1, 1, 2])
        skip, rc = rnd.random() < 0.25, rnd.random() < 0.25
        try:
            exp = bo.extract_se(fq, pat, k, rc, skip)
        except Exception:
            continue
        try:
            ctx = make_ctx(pat, k=k, skip=skip, rc=rc)
        except _lib.BarkitError as e:
            assert e.status in (_lib.BK_E_UNSUPPORTED, _lib.BK_E_PATTERN), (pat, str(e))
            continue
        general += "vm 0 " in _lib.Program(pat, k).dump()
        try:
            out, _ = ctx.extract_text(fq)
        except _lib.BarkitError as e:      # the backtracking budget: refused, not approximated
            assert e.status == _lib.BK_E_UNSUPPORTED, (pat, str(e))
            gave_up += 1
            continue
        finally:
            ctx.close()
        assert out == exp, (pat, k, skip, rc)
        accepted += 1
    assert accepted >= 80 and general >= 60 and gave_up <= accepted // 4, (accepted, general, gave_up)


def test_host_side_mate_falls_back_when_the_host_copy_is_slow():
    """The context measures the host copy of the pattern-less mate against what the bus would need and sends
    both mates to the device once the host side loses twice in a row (here: one host thread).  The bytes do
    not change, the traffic does."""
    p1, p2, _ = CONFIG_PATTERNS["C3"]
    n = 400000
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS["C3"], 0, 2000)
    fq1, fq2 = fq1 * (n // 2000), fq2 * (n // 2000)   # (repeated records: only sizes and timing matter here)
    ctx = make_ctx(p1, p2, k=1, paired=True, host_threads=1)
    seen = []
    for _ in range(5):
        o1, o2, res = ctx.extract_text(fq1, fq2)
        seen.append((zlib.crc32(o1), zlib.crc32(o2), res.n_out, res.h2d_bytes))
    ctx.close()
    assert len({s[:3] for s in seen}) == 1
    assert seen[0][3] < 0.6 * seen[-1][3], seen          # first batches: one mate over the bus; later: both


def test_empty_and_tiny_inputs():
    ctx = make_ctx("^(?P<UMI>[ATGCN]{4})")
    out, res = ctx.extract_text(b"")
    assert out == b"" and res.n_in == 0 and res.n_out == 0
    out, res = ctx.extract_text(b"@x\n\n+\n\n")          # empty read
    assert out == b""
    out, res = ctx.extract_text(b"@x\nACGT\n+\nIIII")     # exactly the pattern, no trailing newline
    assert out == b"@x UMI:ACGT:IIII\n\n+\n\n"
    ctx.close()


def test_long_records_shrink_the_tile():
    rnd = random.Random(5)
    fq = ragged_fastq(rnd, 300, plant=[b"GATTACA"], min_len=1500, max_len=4000)
    pat = "(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})"
    out, _ = run_se(fq, pat, 1)
    assert out == bo.extract_se(fq, pat, 1)


def test_long_records_through_match_kernel():
    """match_kernel sizes a warp's buffer from max_rec_bytes: fewer reads per tile first, then fewer warps per CTA
    (one warp may take most of the SM's shared memory).  Variants, a parametric program and two patterns with a search
    on reads of 1.5-4 kB and of 10-18 kB; a parametric capture can double the record."""
    rnd = random.Random(15)
    for n, lo, hi in ((300, 1500, 4000), (40, 10000, 18000)):
        heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
        fq1 = ragged_fastq(rnd, n, 1, plant=[b"GATTACA"], head_lens=heads, min_len=lo, max_len=hi)
        fq2 = ragged_fastq(rnd, n, 2, plant=[b"ATGCCAT"], head_lens=heads, min_len=lo, max_len=hi)
        for pat in ("(?P<SB>[ATGC]{2,4})gattaca(?P<UMI>[ATGCN]{6})", "^(?P<UMI>[ATGCN]{100,}?)gattaca(?P<CB>[ACGT]*)"):
            if hi > 8000 and "*" in pat:
                # two unbounded captures: the output tile is sized for a record three times its input (par_growth), which
                # together with the four staged input tiles no longer fits the SM for 36 kB records -> BK_E_CAPACITY
                with pytest.raises(_lib.BarkitError) as e:
                    run_se(fq1, pat, 1, extra_flags=_lib.FLAG_VERBOSE)
                assert e.value.status == _lib.BK_E_CAPACITY
                continue
            out, _ = run_se(fq1, pat, 1)
            assert out == bo.extract_se(fq1, pat, 1), (pat, n)
        p1, p2 = "^(?P<UMI>[ATGCN]{12})", "(?P<CB>[ACGT]{6})atgccat"
        # (36 kB records: staging both mates at once does not fit, so neither the paired kernel of
        # BK_FLAG_MATCH_IN_KERNEL nor the paired table launch can run; the default falls back to one single-end
        # launch per mate.  Any extra flag keeps run_pe from also trying the other forms.)
        o1, o2, _ = run_pe(fq1, fq2, p1, p2, 1, extra_flags=_lib.FLAG_HOST_MATE_PINNED if hi > 8000 else 0)
        assert (o1, o2) == bo.extract_pe(fq1, fq2, p1, p2, 1)


@pytest.mark.parametrize("skip", [False, True])
def test_output_tile_holds_records_that_grow(skip):
    """The shared-memory output tile is sized by what a record can grow to: header growth less twice the match length
    when the match is trimmed out of the seq and qual lines (parse.rs:138-148), the whole header growth with -s.
    Patterns whose captures make up the whole match grow every record (3 x (4 + 2 x 20) - 2 x 60 = +12 bytes trimmed,
    +132 with -s); equal-length reads fill every tile to its bound, single-end and as either mate of a pair."""
    p = "^(?P<UMI>[ATGCN]{20})(?P<CB>[ATGCN]{20})(?P<SB>[ATGCN]{20})"
    rnd = random.Random(31)
    n = 3000
    recs1, recs2 = [], []
    for i in range(n):
        L = 61 if i % 97 else 59          # (a few reads too short to match: dropped / kept unchanged)
        s1 = bytes(rnd.choice(b"ACGT") for _ in range(L))
        s2 = bytes(rnd.choice(b"ACGTN") for _ in range(L))
        h = b"r%06d" % i
        recs1.append(b"@" + h + b"\n" + s1 + b"\n+\n" + _qual(L) + b"\n")
        recs2.append(b"@" + h + b"\n" + s2 + b"\n+\n" + _qual(L, 40) + b"\n")
    fq1, fq2 = b"".join(recs1), b"".join(recs2)
    out, res = run_se(fq1, p, 0, skip=skip)
    assert out == bo.extract_se(fq1, p, 0, skip_trimming=skip)
    assert res.n_out == n - len(range(0, n, 97))
    for p1, p2 in ((p, None), (None, p), (p, p)):
        o1, o2, _ = run_pe(fq1, fq2, p1, p2, 0, skip=skip)
        assert (o1, o2) == bo.extract_pe(fq1, fq2, p1, p2, 0, skip_trimming=skip)


def test_output_capacity_error_is_reported():
    ctx = make_ctx("^(?P<UMI>[ATGCN]{12})")
    fq = bksyn.generate(bksyn.CONFIGS["C1"], 1, 0, 1000)
    off, _, maxrec = _lib.tokenize(fq)
    arr = np.zeros(len(fq) + 64, np.uint8)
    arr[:len(fq)] = np.frombuffer(fq, np.uint8)
    m = ctx._mate(arr, off, maxrec)
    out = np.zeros(1000, np.uint8)
    res = _lib.Result()
    rc = _lib.lib.bk_extract_host(ctx.handle, len(off) - 1, C.byref(m), None, out.ctypes.data, out.size,
                                  None, 0, C.byref(res))
    assert rc == _lib.BK_E_CAPACITY
    ctx.close()


# ---- device generator == normative Python generator -----------------------------------------------------------

@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_device_generator_matches_bksyn(name):
    ctx = make_ctx("^(?P<UMI>[ATGCN]{12})")
    cfg = bksyn.CONFIGS[name]
    n, first = 3000, 12345678901
    rl = bksyn.record_len(cfg.read_len)
    d_text = ctx.dev_alloc(n * rl + 64)
    d_off = ctx.dev_alloc((n + 1) * 4)
    for mate in (1, 2):
        ctx.generate_device(lib_syn_config(name), mate, bksyn.DEFAULT_SEED, first, n, d_text, d_off)
        text = ctx.d2h(d_text, n * rl).tobytes()
        off = ctx.d2h(d_off, (n + 1) * 4, np.uint32)
        assert text == bksyn.generate(cfg, mate, first, n)
        assert (off == np.arange(n + 1, dtype=np.uint64) * rl).all()
    ctx.dev_free(d_text)
    ctx.dev_free(d_off)
    ctx.close()


# ---- device-resident path, match table, full-size properties --------------------------------------------------

def _device_batch(ctx, name, first, n):
    cfg = bksyn.CONFIGS[name]
    rl = bksyn.record_len(cfg.read_len)
    mates = []
    for mate in (1, 2):
        d_text = ctx.dev_alloc(n * rl + 64)
        d_off = ctx.dev_alloc((n + 1) * 4)
        ctx.generate_device(lib_syn_config(name), mate, bksyn.DEFAULT_SEED, first, n, d_text, d_off)
        m = _lib.MateIn()
        m.text, m.rec_off, m.text_len, m.max_rec_bytes = d_text, d_off, n * rl, rl
        mates.append(m)
    return mates, rl


def test_device_resident_matches_oracle_and_match_table():
    name, n, first = "C4", 5000, 777
    p1, p2, _ = CONFIG_PATTERNS[name]
    ctx = make_ctx(p1, p2, k=2, paired=True)
    mates, rl = _device_batch(ctx, name, first, n)
    cap = n * (rl + 120)
    d_o1, d_o2 = ctx.dev_alloc(cap), ctx.dev_alloc(cap)
    d_m1, d_m2 = ctx.dev_alloc(n * 4), ctx.dev_alloc(n * 4)
    res = ctx.extract_device(n, mates[0], mates[1], d_o1, cap, d_o2, cap, d_m1, d_m2, iters=2)
    fq1, fq2 = bksyn.generate_pair(bksyn.CONFIGS[name], first, n)
    e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, 2)
    assert ctx.d2h(d_o1, res.out1_len).tobytes() == e1
    assert ctx.d2h(d_o2, res.out2_len).tobytes() == e2
    m1 = ctx.d2h(d_m1, n * 4, np.int32)
    m2 = ctx.d2h(d_m2, n * 4, np.int32)
    r1, r2 = bo.BarcodeRegex(p1, 2), bo.BarcodeRegex(p2, 2)
    for i, (a, b) in enumerate(zip(bo.read_fastq(fq1), bo.read_fastq(fq2))):
        ma, mb = r1.regex.search(a.seq), r2.regex.search(b.seq)
        assert m1[i] == (ma.start() if ma else -1)
        assert m2[i] == (mb.start() if mb else -1)
    assert res.launches == 4 and res.kernel_ms > 0   # per iteration: match_kernel on R2, the paired launch in table form
    ctx.close()


def test_full_size_properties_c3():
    """2 M pairs (device-generated): size-independent properties instead of a byte diff.
    (1) out lengths follow the closed form from n_match; (2) kept mate-2 records are byte-identical
    to the input records, in order; (3) running the batch in two halves concatenates to the same
    bytes (ordered compaction is tile-size independent)."""
    name, n, first = "C3", 2_000_000, 5_000_000
    p1, p2, _ = CONFIG_PATTERNS[name]
    ctx = make_ctx(p1, p2, k=1, paired=True)
    mates, rl = _device_batch(ctx, name, first, n)
    cap = n * (rl + 70)
    d_o1, d_o2, d_m1 = ctx.dev_alloc(cap), ctx.dev_alloc(cap), ctx.dev_alloc(n * 4)
    res = ctx.extract_device(n, mates[0], mates[1], d_o1, cap, d_o2, cap, d_m1, 0)
    assert res.n_out == res.n_match1 and res.n_match2 == 0
    assert 0.88 < res.n_out / n < 0.92
    assert res.out1_len == res.n_out * (rl + 67 - 2 * 35)       # +header text, -35 from seq and qual
    assert res.out2_len == res.n_out * rl
    m1 = ctx.d2h(d_m1, n * 4, np.int32)
    keep = m1 >= 0
    assert set(np.unique(m1)) <= {-1, 0}
    assert int(keep.sum()) == res.n_out
    in2 = ctx.d2h(mates[1].text, n * rl).reshape(n, rl)
    out2 = ctx.d2h(d_o2, res.out2_len).reshape(-1, rl)
    assert (out2 == in2[keep]).all()
    whole1 = zlib.crc32(ctx.d2h(d_o1, res.out1_len).tobytes())
    # halves
    h = n // 2 + 13
    crc, total = 0, 0
    for a, b in ((0, h), (h, n)):
        ma, mb = _lib.MateIn(), _lib.MateIn()
        for src, dst in ((mates[0], ma), (mates[1], mb)):
            dst.text = src.text + a * rl
            dst.rec_off = ctx.dev_alloc((b - a + 1) * 4)
            ctx.h2d(dst.rec_off, (np.arange(b - a + 1, dtype=np.uint64) * rl).astype(np.uint32))
            dst.text_len, dst.max_rec_bytes = (b - a) * rl, rl
        # text pointers must stay 16-byte aligned: rl*h is not in general, so realign through a copy
        if (a * rl) % 16:
            for dst in (ma, mb):
                tmp = ctx.dev_alloc((b - a) * rl + 64)
                ctx.h2d(tmp, ctx.d2h(dst.text, (b - a) * rl))
                dst.text = tmp
        r = ctx.extract_device(b - a, ma, mb, d_o1, cap, d_o2, cap)
        crc = zlib.crc32(ctx.d2h(d_o1, r.out1_len).tobytes(), crc)
        total += r.out1_len
    assert total == res.out1_len and crc == whole1
    ctx.close()


# ---- bytes >= 0x80 (declared divergence) ----------------------------------------------------

def _fq(seqs):
    return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs))


def test_invalid_utf8_bytes_never_match():
    """regex::bytes::Regex runs in Unicode mode (pattern.rs:6,182): a byte >= 0x80 outside a well-formed UTF-8
    sequence matches neither `.` (a fuzzy run's wildcard slot) nor a class.  The oracle states that (its `.` is the
    UTF-8 scalar alternation); the kernel's rule -- bytes >= 0x80 never match -- gives the same bytes here."""
    rnd = random.Random(2)
    pat = "^(?P<UMI>[ATGCN]{4})atgc(?P<CB>[ATGCN]{3})"
    seqs = []
    for i in range(4000):
        s = bytearray(b"ACGT" + b"ATGC" + b"TTT" + bytes(rnd.choice(b"ACGT") for _ in range(rnd.randrange(0, 40))))
        for _ in range(rnd.choice([0, 1, 1, 2])):
            s[rnd.randrange(0, min(len(s), 14))] = rnd.choice([0x80, 0xA9, 0xBF, 0xC0, 0xC1, 0xF5, 0xFF, 0xFE, ord("N"), ord("x")])
        # (continuation bytes and the never-valid lead bytes only: no well-formed multi-byte sequence can arise)
        seqs.append(bytes(s))
    fq = _fq(seqs)
    for k in (0, 1, 2):
        out, res = run_se(fq, pat, k=k)
        assert out == bo.extract_se(fq, pat, k)
        assert 0 < res.n_out < len(seqs)
    # unanchored search and the reverse-complement retry go through other code
    upat = "(?P<UMI>[ATGCN]{4})atgc"
    assert run_se(fq, upat, k=1)[0] == bo.extract_se(fq, upat, 1)
    # -r: the reverse-complement match is found on complemented text (every byte a letter) and its offsets applied to
    # the forward read, so a capture CAN hold such bytes: from_utf8 fails and the read is dropped (parse.rs:166)
    out_rc, res_rc = run_se(fq, pat, k=1, rc=True)
    assert out_rc == bo.extract_se(fq, pat, 1, rc_barcodes=True)
    # ... and kept when the captured bytes are a well-formed sequence
    seqs2 = [b"AAA" + b"GCAT" + b"ACGT", b"AAA" + b"GCAT" + b"A\xc3\xa9T", b"AAA" + b"GCAT" + b"AC\xffT", b"AAA" + b"GCAT" + b"AC\xc3"]
    fq2 = _fq(seqs2)
    out2, res2 = run_se(fq2, pat, k=0, rc=True)
    assert out2 == bo.extract_se(fq2, pat, 0, rc_barcodes=True) and res2.n_out == 2


def test_well_formed_multibyte_sequences_are_the_declared_divergence():
    """A WELL-FORMED multi-byte sequence under a wildcard slot is one `.` for the regex crate (the match grows by the
    extra bytes); here no byte >= 0x80 ever matches, so such a read is simply unmatched.  Sequence lines of FASTQ are
    ASCII; this pins the behaviour that DESIGN.md declares instead of leaving it unobserved."""
    pat = "^(?P<UMI>[ATGCN]{4})atgc"
    plain, odd = b"ACGTATGCAAAA", b"ACGTA\xc3\xa9GCAAAA"          # U+00E9 in the wildcard slot of A.GC
    fq = _fq([plain, odd, plain])
    ref = bo.extract_se(fq, pat, 1)
    assert ref.count(b"\n") == 12                                 # the reference keeps all three reads
    out, res = run_se(fq, pat, k=1)
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate([plain, odd, plain])]
    assert res.n_out == 2 and out == bo.extract_se(recs[0] + recs[2], pat, 1)
