"""bk::gather_kept (bk_fastq.cc): the host-side assembly of a pattern-less mate's output in paired mode -- the kept
records of the caller's input buffer, in input order, in the writer's framing (run.rs:149-160, fastq.rs:259-263).
It runs inside bk_wait on every default-mode batch of a one-pattern pair, i.e. on the end-to-end path the bench
headlines; no GPU is involved, so it is pinned here against the oracle's reader and writer."""
import ctypes as C
import random

import numpy as np
import pytest

from barkit_b200 import _lib
from oracle import barkit_oracle as bo
from tests.util import ragged_fastq

gather = getattr(_lib.lib, "_ZN2bk11gather_keptEPKhPKjjPKiPhmi")
gather.restype = C.c_uint64
gather.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]


def run(fq: bytes, keep: np.ndarray, threads: int, cap=None):
    off, consumed, _ = _lib.tokenize(fq)
    assert consumed == len(fq) and len(off) == len(keep) + 1
    text = np.frombuffer(fq + b"\0" * 64, np.uint8).copy()
    out = np.full((len(fq) + 80) if cap is None else max(cap, 1), 0xEE, np.uint8)
    n = gather(text.ctypes.data, off.ctypes.data, len(keep), keep.ctypes.data, out.ctypes.data,
               len(fq) + 16 if cap is None else cap, threads)
    return n, out


def expected(fq: bytes, keep: np.ndarray) -> bytes:
    recs = bo.read_fastq(fq)
    return bo.write_fastq([r for r, k in zip(recs, keep) if k >= 0])


@pytest.mark.parametrize("crlf,plus,final_nl", [(False, False, True), (True, False, True), (False, True, True),
                                                (False, False, False), (True, True, False)])
@pytest.mark.parametrize("threads", [1, 5])
def test_kept_records_in_writer_framing(crlf, plus, final_nl, threads):
    rnd = random.Random(11 + 2 * crlf + plus)
    fq = ragged_fastq(rnd, 700, mate=2, crlf=crlf, plus_comment=plus, final_newline=final_nl, max_len=300)
    for p_keep in (0.0, 0.5, 0.9, 1.0):
        keep = np.array([rnd.randrange(0, 40) if rnd.random() < p_keep else -1 for _ in range(700)], np.int32)
        n, out = run(fq, keep, threads)
        want = expected(fq, keep)
        assert n == len(want)
        assert bytes(out[:n]) == want
        assert (out[n + 64:] == 0xEE).all()      # nothing written far beyond the output (the closing-up moves left)


def test_many_threads_long_runs_and_every_alignment():
    """Large enough for the threaded path (>= 50 000 records) with long runs of kept records: the streaming copy's
    head / aligned middle / tail at every destination alignment, parts closed up across threads."""
    rnd = random.Random(5)
    recs = []
    for i in range(60000):
        L = 150 if i % 1000 else rnd.randint(0, 40)           # mostly equal lengths, a few odd ones shift the alignment
        s = bytes(rnd.choice(b"ACGT") for _ in range(8)) * (L // 8) + b"A" * (L % 8)
        recs.append(b"@r%d 2:N:0:1\n" % i + s + b"\n+\n" + b"I" * L + b"\n")
    fq = b"".join(recs)
    keep = np.array([-1 if (i % 10 == 3 or 20000 <= i < 20100) else 7 for i in range(60000)], np.int32)
    want = expected(fq, keep)
    for threads in (1, 3, 8):
        n, out = run(fq, keep, threads)
        assert n == len(want) and bytes(out[:n]) == want


def test_capacity_is_checked():
    rnd = random.Random(2)
    fq = ragged_fastq(rnd, 50, mate=2)
    keep = np.zeros(50, np.int32)
    n, _ = run(fq, keep, 1, cap=len(fq) // 2)
    assert n == 2 ** 64 - 1
