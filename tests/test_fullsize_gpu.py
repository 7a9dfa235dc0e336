"""Full-size byte parity ("no full-size byte parity").  Every BASELINE.json config at 2 M units:
the CUDA path (device-resident form, and host buffers through bk_submit / bk_wait with the pattern-less mate on
the host and on the device) against oracle/cpu_extract.c, byte for byte; a C5-style run of counter-indexed
batches dealt over every visible GPU with a sample of batches byte-diffed and a CRC chain of all output compared
with the 1-GPU run; and the `barkit` CLI with --gpus 1 against --gpus N.  Look-back depth, 64-bit prefixes and
the block-of-32 roll-over (thousands of tiles) are exercised by size alone."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from barkit_b200 import _lib
from oracle import bksyn, cport
from tests.util import CONFIG_PATTERNS, lib_syn_config

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "barkit_b200", "barkit")
N_FULL = 2_000_000
SEED = 20240902
RL = bksyn.record_len(150)
THREADS = max(1, min(len(os.sched_getaffinity(0)), 32))


def _ctx(name, k, device=0, extra_flags=0):
    p1, p2, paired = CONFIG_PATTERNS[name]
    return _lib.Context(_lib.Program(p1, k) if p1 else None, _lib.Program(p2, k) if p2 else None, paired=paired,
                        device=device, extra_flags=extra_flags), paired


def _device_batch(ctx, name, first_idx, n, paired):
    cfg = lib_syn_config(name)
    mates = []
    for mate in ((1, 2) if paired else (1,)):
        d_text, d_off = ctx.dev_alloc(n * RL + 64), ctx.dev_alloc((n + 1) * 4)
        ctx.generate_device(cfg, mate, SEED, first_idx, n, d_text, d_off)
        m = _lib.MateIn()
        m.text, m.rec_off, m.text_len, m.max_rec_bytes = d_text, d_off, n * RL, RL
        mates.append(m)
    return mates


def _free(ctx, mates, *ptrs):
    for m in mates:
        ctx.dev_free(m.text)
        ctx.dev_free(m.rec_off)
    for p in ptrs:
        if p:
            ctx.dev_free(p)


def _oracle(name, k, texts, n):
    p1, p2, paired = CONFIG_PATTERNS[name]
    off = (np.arange(n + 1, dtype=np.uint64) * RL).astype(np.uint32)
    r = cport.extract(texts[0], texts[1] if paired else None, p1, p2, k, threads=THREADS, off1=off,
                      off2=off if paired else None)
    return r


def _first_diff(a, b):
    a, b = np.frombuffer(a, np.uint8), np.frombuffer(b, np.uint8)
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if d.size else (n if a.size != b.size else -1)


def _assert_same(got, want, what):
    if got != want:
        i = _first_diff(got, want)
        raise AssertionError("%s: %d vs %d bytes, first difference at byte %d: %r / %r" %
                             (what, len(got), len(want), i, got[max(0, i - 40):i + 40], want[max(0, i - 40):i + 40]))


@pytest.mark.parametrize("name,k", [("C1", 1), ("C2", 1), ("C3", 1), ("C4", 0), ("C4", 1), ("C4", 2)])
def test_full_size_byte_parity_device_resident(name, k):
    ctx, paired = _ctx(name, k)
    n = N_FULL
    mates = _device_batch(ctx, name, 0, n, paired)
    cap = n * (RL + 120)
    d_o1, d_o2 = ctx.dev_alloc(cap), (ctx.dev_alloc(cap) if paired else 0)
    res = ctx.extract_device(n, mates[0], mates[1] if paired else None, d_o1, cap, d_o2, cap if paired else 0)
    texts = [ctx.d2h(m.text, n * RL) for m in mates]
    want = _oracle(name, k, texts, n)
    assert res.n_out == want["n_out"] and res.n_in == n
    _assert_same(ctx.d2h(d_o1, res.out1_len).tobytes(), want["out1"], "%s k=%d out1" % (name, k))
    if paired:
        _assert_same(ctx.d2h(d_o2, res.out2_len).tobytes(), want["out2"], "%s k=%d out2" % (name, k))
    _free(ctx, mates, d_o1, d_o2)
    ctx.close()


@pytest.mark.parametrize("name,k,flags", [("C1", 1, 0), ("C2", 1, 0), ("C2", 1, _lib.FLAG_DEVICE_BOTH_MATES),
                                          ("C3", 1, 0), ("C3", 1, _lib.FLAG_DEVICE_BOTH_MATES),
                                          ("C3", 1, _lib.FLAG_HOST_MATE_PINNED), ("C4", 1, 0)])
def test_full_size_byte_parity_host_buffers(name, k, flags):
    """bk_submit / bk_wait on pinned host buffers, both slots in flight, 2 x 1 M units."""
    ctx, paired = _ctx(name, k, extra_flags=flags)
    n = N_FULL // 2
    nb = n * RL
    res = _lib.Result()
    got = {0: [], 1: []}
    want_all = []
    bufs = []
    for b in range(2):
        mates = _device_batch(ctx, name, b * n, n, paired)
        texts = [ctx.d2h(m.text, nb + 64) for m in mates]
        off = ctx.d2h(mates[0].rec_off, (n + 1) * 4, np.uint32)
        want_all.append(_oracle(name, k, [t[:nb] for t in texts], n))
        ins = []
        for t in texts:
            h = _lib.MateIn()
            h.text, h.rec_off, h.text_len, h.max_rec_bytes = t.ctypes.data, off.ctypes.data, nb, RL
            ins.append(h)
        outs = [np.zeros(int(ctx.out_bound(i + 1, n, nb)) + 16, np.uint8) for i in range(len(mates))]
        bufs.append((texts, off, ins, outs))
        _free(ctx, mates)
        rc = _lib.lib.bk_submit(ctx.handle, b, n, C.byref(ins[0]), C.byref(ins[1]) if paired else None)
        assert rc == 0, _lib.lib.bk_last_error(ctx.handle)
    for b in range(2):
        _, _, _, outs = bufs[b]
        rc = _lib.lib.bk_wait(ctx.handle, b, outs[0].ctypes.data, outs[0].size, outs[1].ctypes.data if paired else None,
                              outs[1].size if paired else 0, C.byref(res))
        assert rc == 0, _lib.lib.bk_last_error(ctx.handle)
        want = want_all[b]
        assert res.n_out == want["n_out"]
        _assert_same(outs[0][:res.out1_len].tobytes(), want["out1"], "%s batch %d out1" % (name, b))
        if paired:
            _assert_same(outs[1][:res.out2_len].tobytes(), want["out2"], "%s batch %d out2" % (name, b))
        if name in ("C2", "C3"):
            one_mate = nb + (n + 1) * 4
            assert res.h2d_bytes == (2 * one_mate if flags & _lib.FLAG_DEVICE_BOTH_MATES else one_mate)
    ctx.close()


def test_c5_style_batches_over_all_gpus():
    """SURVEY.md section 8d C5, scaled to a test: 24 batches of 500 k pairs by counter-based index, dealt round-robin
    over every visible GPU (one context each; no collective), every 6th batch byte-diffed against the C oracle and a
    CRC chain over all output, in batch order, compared with the same batches run on GPU 0 alone."""
    name, k, n, nbatches = "C3", 1, 500_000, 24
    ndev = _lib.device_count()
    cap = n * (RL + 120)

    def run(devices):
        ctxs = [_ctx(name, k, device=d, extra_flags=0)[0] for d in devices]
        outs = [(c.dev_alloc(cap), c.dev_alloc(cap)) for c in ctxs]
        crc1 = crc2 = 0
        kept = 0
        for b in range(nbatches):
            i = b % len(ctxs)
            ctx = ctxs[i]
            mates = _device_batch(ctx, name, b * n, n, True)
            res = ctx.extract_device(n, mates[0], mates[1], outs[i][0], cap, outs[i][1], cap)
            o1, o2 = ctx.d2h(outs[i][0], res.out1_len).tobytes(), ctx.d2h(outs[i][1], res.out2_len).tobytes()
            if b % 6 == 0:
                want = _oracle(name, k, [ctx.d2h(m.text, n * RL) for m in mates], n)
                _assert_same(o1, want["out1"], "batch %d out1" % b)
                _assert_same(o2, want["out2"], "batch %d out2" % b)
            crc1, crc2, kept = zlib.crc32(o1, crc1), zlib.crc32(o2, crc2), kept + res.n_out
            _free(ctx, mates)
        for c in ctxs:
            c.close()
        return crc1, crc2, kept

    single = run([0])
    assert 0.88 < single[2] / (n * nbatches) < 0.92
    assert run(list(range(ndev))) == single        # (on a 1-GPU box: the same run twice must be bit-identical)


def _write_pair(tmp_path, name, n, first=0):
    t1 = cport.generate(bksyn.CONFIGS[name], 1, first, n, SEED, THREADS)
    t2 = cport.generate(bksyn.CONFIGS[name], 2, first, n, SEED, THREADS)
    (tmp_path / "r1.fq").write_bytes(t1.tobytes())
    (tmp_path / "r2.fq").write_bytes(t2.tobytes())
    return t1, t2


@pytest.mark.parametrize("name,k", [("C3", 1), ("C4", 1)])
def test_cli_gpus_1_and_n_write_identical_files(tmp_path, name, k):
    """bk_run's multi-GPU path (dispatcher, per-GPU workers, ordered writer): -m 1 gives ~190 batches of 1 MiB, so
    every worker takes many and completion order is shuffled; --gpus 1 and --gpus N must write the same bytes, equal
    to the oracle's.  On a 1-GPU box the second leg repeats --gpus 1 through the automatic device choice."""
    p1, p2, _ = CONFIG_PATTERNS[name]
    n = 300_000
    t1, t2 = _write_pair(tmp_path, name, n)
    want = cport.extract(t1, t2, p1, p2, k, threads=THREADS)
    ndev = _lib.device_count()
    outs = []
    for tag, gpus in (("a", ["--gpus", "1"]), ("b", ["--gpus", str(ndev)] if ndev > 1 else [])):
        args = [CLI, "-m", "1", "--quiet", "-f", "extract", "-1", str(tmp_path / "r1.fq"), "-2", str(tmp_path / "r2.fq"),
                "-o", str(tmp_path / ("o1%s.fq" % tag)), "-O", str(tmp_path / ("o2%s.fq" % tag)), "-e", str(k)] + gpus
        if p1:
            args += ["-p", p1]
        if p2:
            args += ["-P", p2]
        r = subprocess.run(args, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(((tmp_path / ("o1%s.fq" % tag)).read_bytes(), (tmp_path / ("o2%s.fq" % tag)).read_bytes()))
    _assert_same(outs[0][0], want["out1"], "--gpus 1 out1")
    _assert_same(outs[0][1], want["out2"], "--gpus 1 out2")
    assert outs[1] == outs[0]
