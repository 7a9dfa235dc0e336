"""bk_tokenize_device (record boundaries found on the device) against the host
tokenizer bk_tokenize, which follows seq_io's reader (fastq.rs:166-172).  Needs a B200."""
import ctypes as C
import random

import numpy as np
import pytest

from barkit_b200 import _lib
from oracle import barkit_oracle as bo
from oracle import bksyn
from tests.util import ragged_fastq

pytestmark = pytest.mark.gpu


def host_tokenize(text, final=True, cap=None):
    arr = np.frombuffer(text, np.uint8) if text else np.zeros(1, np.uint8)
    if cap is None:
        cap = int(np.count_nonzero(arr == 10)) // 4 + 2
    off = np.zeros(cap + 1, np.uint32)
    n, consumed, maxrec = C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
    rc = _lib.lib.bk_tokenize(arr.ctypes.data, len(text), 1 if final else 0, off.ctypes.data, cap,
                              C.byref(n), C.byref(consumed), C.byref(maxrec))
    return rc, off[:n.value + 1].copy(), consumed.value, maxrec.value


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(_lib.Program("^(?P<UMI>[ATGCN]{4})", 1))
    yield c
    c.close()


def same(ctx, text, final=True, cap=None, ends_clean=True):
    rh, oh, ch, mh = host_tokenize(text, final, cap)
    rd, od, cd, md = ctx.tokenize_device(text, final, cap)
    assert rd == rh, (rd, rh)
    assert len(od) == len(oh), (len(od), len(oh))
    assert (od[:-1] == oh[:-1]).all()
    assert cd == ch
    if rh == _lib.BK_OK:
        assert md == mh or not ends_clean
        if ends_clean:
            assert od[-1] == oh[-1]
    return rd, od


def test_synthetic_blocks(ctx):
    for n in (1, 7, 4096 // 329 + 1, 5000):
        fq = bksyn.generate(bksyn.CONFIGS["C3"], 1, 0, n)
        rc, off = same(ctx, fq)
        assert rc == _lib.BK_OK and len(off) == n + 1 and off[-1] == len(fq)


@pytest.mark.parametrize("seed", range(6))
def test_ragged_text(ctx, seed):
    rnd = random.Random(seed)
    fq = ragged_fastq(rnd, 3000 + seed, crlf=(seed == 2), plus_comment=(seed % 2 == 1), final_newline=(seed != 3))
    same(ctx, fq)
    same(ctx, fq, cap=100)                 # the caller's record limit
    same(ctx, fq[:len(fq) * 2 // 3], final=False)   # a chunk that ends inside a record


def test_edges(ctx):
    rec = b"@r\nACGT\n+\nIIII\n"
    assert same(ctx, b"")[0] == _lib.BK_OK
    same(ctx, rec)
    same(ctx, rec[:-1])                      # last line without EOL
    same(ctx, rec * 3 + b"\n\n\r\n", ends_clean=False)   # blank lines after the last record
    same(ctx, b"\n\n", ends_clean=False)
    same(ctx, rec * 2 + rec[:9])             # truncated final record
    same(ctx, rec * 2 + rec[:9], final=False)
    same(ctx, rec + b"r\nACGT\n+\nIIII\n" + rec)      # no '@'
    same(ctx, rec + b"@r\nACGT\n-\nIIII\n" + rec)     # no '+'
    same(ctx, rec + b"@r\nACGT\n+\nIII\n" + rec)      # seq / qual lengths differ
    same(ctx, rec + b"\n" + rec)             # blank line between records


def test_extract_from_device_tokenized_text(ctx):
    """end to end: offsets found on the device feed the extract kernel; output equals the oracle's"""
    rnd = random.Random(11)
    fq = ragged_fastq(rnd, 2000, plant=[b"ACGT"], plus_comment=True)
    rc, off, _, maxrec = ctx.tokenize_device(fq)
    assert rc == _lib.BK_OK
    arr = np.zeros(len(fq) + 64, np.uint8)
    arr[:len(fq)] = np.frombuffer(fq, np.uint8)
    m = ctx._mate(arr, off, maxrec)
    cap = len(fq) + 40 * (len(off) - 1) + 64
    out = np.zeros(cap, np.uint8)
    res = _lib.Result()
    assert _lib.lib.bk_extract_host(ctx.handle, len(off) - 1, C.byref(m), None, out.ctypes.data, cap, None, 0,
                                    C.byref(res)) == _lib.BK_OK
    assert out[:res.out1_len].tobytes() == bo.extract_se(fq, "^(?P<UMI>[ATGCN]{4})", 1)


# ---- against the oracle's reader (oracle/barkit_oracle.py: read_fastq = seq_io's field definitions) ----------------

@pytest.mark.parametrize("seed", range(5))
def test_offsets_cut_the_oracles_records(ctx, seed):
    """The device's record offsets against the ORACLE's tokeniser, not the product's own host one: the text between
    consecutive offsets must be exactly one record for read_fastq, and the records in order must be read_fastq's of
    the whole text."""
    rnd = random.Random(100 + seed)
    fq = ragged_fastq(rnd, 2500 + 7 * seed, crlf=(seed == 1), plus_comment=(seed in (2, 3)), final_newline=(seed != 4))
    want = bo.read_fastq(fq)
    rc, off, consumed, maxrec = ctx.tokenize_device(fq, True, None)
    assert rc == _lib.BK_OK and len(off) == len(want) + 1 and off[0] == 0 and consumed == len(fq)
    longest = 0
    for r, rec in enumerate(want):
        piece = fq[off[r]:off[r + 1]] if r + 1 < len(want) else fq[off[r]:]
        got = bo.read_fastq(piece)
        assert len(got) == 1 and got[0] == rec, r
        longest = max(longest, len(piece))
    assert maxrec == longest


def test_malformed_text_is_refused_where_the_oracle_refuses_it(ctx):
    rec = b"@r\nACGT\n+\nIIII\n"
    for bad in (rec + b"r\nACGT\n+\nIIII\n" + rec, rec + b"@r\nACGT\n-\nIIII\n" + rec, rec * 2 + rec[:9]):
        with pytest.raises(ValueError):
            bo.read_fastq(bad)
        assert ctx.tokenize_device(bad, True, None)[0] == _lib.BK_E_FORMAT
