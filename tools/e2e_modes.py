"""End-to-end rate of bk_submit / bk_wait (offsets from the host) against bk_submit_text (raw text, device tokenizer), both
mates over the bus, pinned host buffers, two slots in flight.   python tools/e2e_modes.py [pairs per step] [steps]
BK_E2E_FLAGS=<int> replaces the context flags (0: the library's default, the pattern-less mate assembled on the host);
BK_DEBUG=1 adds the library's per-batch timing lines."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
from tests.util import CONFIG_PATTERNS, lib_syn_config
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rl = 329; nb = Q * rl
p1 = CONFIG_PATTERNS["C3"][0]
ctx = _lib.Context(_lib.Program(p1, 1), None, paired=True, extra_flags=int(os.environ.get("BK_E2E_FLAGS", _lib.FLAG_DEVICE_BOTH_MATES)) | (_lib.FLAG_VERBOSE if os.environ.get("BK_DEBUG") else 0))
h = {}
for name, size in (("t1", nb + 64), ("t2", nb + 64), ("off", (Q + 1) * 4), ("o1a", ctx.out_bound(1, Q, nb)), ("o2a", ctx.out_bound(2, Q, nb)),
                   ("o1b", ctx.out_bound(1, Q, nb)), ("o2b", ctx.out_bound(2, Q, nb))):
    h[name] = (_lib.lib.bk_host_alloc(size), size)
for mate, t in ((1, "t1"), (2, "t2")):
    dt = ctx.dev_alloc(nb + 64); do = ctx.dev_alloc((Q + 1) * 4)
    ctx.generate_device(lib_syn_config("C3"), mate, 20240902, 0, Q, dt, do)
    _lib.lib.bk_memcpy_d2h(ctx.handle, h[t][0], dt, nb)
    _lib.lib.bk_memcpy_d2h(ctx.handle, h["off"][0], do, (Q + 1) * 4)
    ctx.dev_free(dt); ctx.dev_free(do)
m1, m2 = _lib.MateIn(), _lib.MateIn()
for m, t in ((m1, "t1"), (m2, "t2")):
    m.text, m.rec_off, m.text_len, m.max_rec_bytes = h[t][0], h["off"][0], nb, rl
res = _lib.Result()
outs = [("o1a", "o2a"), ("o1b", "o2b")]
def run(raw, k):
    def submit(s):
        rc = (_lib.lib.bk_submit_text(ctx.handle, s, Q, h["t1"][0], nb, h["t2"][0], nb, rl) if raw
              else _lib.lib.bk_submit(ctx.handle, s, Q, C.byref(m1), C.byref(m2)))
        assert rc == 0, _lib.lib.bk_last_error(ctx.handle)
    def wait(s):
        a, b = outs[s]
        rc = _lib.lib.bk_wait(ctx.handle, s, h[a][0], h[a][1], h[b][0], h[b][1], C.byref(res))
        assert rc == 0, _lib.lib.bk_last_error(ctx.handle)
    submit(0)
    for s in range(k):
        if s + 1 < k: submit((s + 1) & 1)
        wait(s & 1)
for raw in (False, True, False, True):
    run(raw, 3)
    t0 = time.perf_counter(); run(raw, K); dt = time.perf_counter() - t0
    print("%-28s %6.1f M pairs/s  %.2f ms per step of %d pairs; out %d + %d bytes, launches %d, kernel %.3f ms" %
          ("bk_submit_text (raw)" if raw else "bk_submit (host offsets)", K * Q / dt / 1e6, dt / K * 1e3, Q, res.out1_len, res.out2_len, res.launches, res.kernel_ms), flush=True)
