"""Pattern fuzzing beyond tests/test_extract_gpu.py::test_random_patterns_against_the_oracle: random patterns over the
lowered grammar on both mates of ragged pairs (CR LF, '+' comments, -r, -s, every launch form), compared byte for byte
with the oracle.  Prints every mismatch; `tried N accepted M` at the end (the rest was refused by the compiler or is
not valid for the oracle's engine).   python tools/fuzz_patterns.py"""
import random, sys, os
sys.path.insert(0, os.getcwd())
import tests.test_extract_gpu as t
from barkit_b200 import _lib
from oracle import barkit_oracle as bo
from tests.util import ragged_fastq
tot=acc=0
for seed in range(1, 9):
    rnd = random.Random(seed * 7919)
    n = 500
    heads = [rnd.choice([3, 23, 60]) for _ in range(n)]
    fq1 = ragged_fastq(rnd, n, 1, plant=[b"GATTACA", b"ACGT", b"ATGC", b"ACACGG", b"TTATTA"], head_lens=heads, min_len=0, max_len=60, crlf=(seed % 3 == 0))
    fq2 = ragged_fastq(rnd, n, 2, plant=[b"GATTACA", b"ACGT", b"ATGC", b"ACACGG", b"TTATTA"], head_lens=heads, min_len=0, max_len=60, plus_comment=True)
    for it in range(300):
        p1 = t._random_pattern(rnd); p2 = t._random_pattern(rnd) if rnd.random() < 0.6 else None
        if rnd.random() < 0.15: p1, p2 = None, (p2 or t._random_pattern(rnd))
        k = rnd.choice([0, 1, 1, 2]); skip, rc = rnd.random() < 0.25, rnd.random() < 0.25
        tot += 1
        try: e1, e2 = bo.extract_pe(fq1, fq2, p1, p2, k, rc_barcodes=rc, skip_trimming=skip)
        except Exception: continue
        try: ctx = t.make_ctx(p1, p2, k=k, paired=True, skip=skip, rc=rc, extra_flags=rnd.choice([0, 0, _lib.FLAG_DEVICE_BOTH_MATES, _lib.FLAG_SPLIT_TWO_PATTERNS, _lib.FLAG_MATCH_IN_KERNEL]))
        except _lib.BarkitError: continue
        try:
            o1, o2, _ = ctx.extract_text(fq1, fq2)
        except _lib.BarkitError as ex:
            print("RUNTIME ERROR", seed, it, repr(p1), repr(p2), k, skip, rc, ex); ctx.close(); continue
        ctx.close()
        acc += 1
        if (o1, o2) != (e1, e2):
            print("MISMATCH", seed, it, repr(p1), repr(p2), k, skip, rc, flush=True)
print("tried", tot, "accepted", acc)
