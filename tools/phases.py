"""Pipeline event trace of one extract launch (needs a BK_PROFILE_PHASES build:
  BK_PROFILE_PHASES=1 python -m barkit_b200.build --force).  Prints percentiles of the intervals between
the per-tile events, in SM cycles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from barkit_b200 import _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 30          # records per tile the launch plan picks (BK_FLAG_VERBOSE prints it)
NAME = sys.argv[3] if len(sys.argv) > 3 else "C3"          # BASELINE.json config
K = int(sys.argv[4]) if len(sys.argv) > 4 else 1
from tests.util import CONFIG_PATTERNS, lib_syn_config
p1, p2, _paired = CONFIG_PATTERNS[NAME]
assert _paired, "the trace is wired for paired-end launches"
ctx = _lib.Context(_lib.Program(p1, K) if p1 else None, _lib.Program(p2, K) if p2 else None, paired=True)
cfg = lib_syn_config(NAME); rl = 329
ms = []
for mate in (1, 2):
    dt = ctx.dev_alloc(P * rl + 64); do = ctx.dev_alloc((P + 1) * 4)
    ctx.generate_device(cfg, mate, 20240902, 0, P, dt, do)
    m = _lib.MateIn(); m.text, m.rec_off, m.text_len, m.max_rec_bytes = dt, do, P * rl, rl; ms.append(m)
cap = P * (rl + 80); o1 = ctx.dev_alloc(cap); o2 = ctx.dev_alloc(cap)
for i in range(3): r = ctx.extract_device(P, ms[0], ms[1], o1, cap, o2, cap)
ntiles = (P + R - 1) // R
lo, n = ntiles // 4, ntiles // 2                            # the steady-state middle of the launch
ev = np.zeros((n, 16), np.uint64)
rc = _lib.lib.bk_debug_events(ctx.handle, ntiles, lo, n, ev.ctypes.data)
assert rc == 0, rc
ev = ev.astype(np.int64)
print("kernel_ms", r.kernel_ms, "Mpairs/s", P / r.kernel_ms / 1e3, "tiles", ntiles)
def iv(a, b, label):
    d = ev[:, b] - ev[:, a]
    d = d[(ev[:, a] > 0) & (ev[:, b] > 0)]
    if len(d) == 0:
        return
    print("%-44s mean %7.0f  p10 %7.0f  p50 %7.0f  p90 %7.0f  p99 %7.0f" % (label, d.mean(), *np.percentile(d, [10, 50, 90, 99])))
names = ["slot free", "claimed", "tma issued", "F0 data landed", "F0 parsed", "F0 matched", "F0 published", "F0 handed over",
         "S start", "S ready", "B00 start", "B00 emit done", "B0 all parts done", "B0 stored", "F1 published", "B03 emit done"]
iv(0, 1, "slot free -> claimed (atomic)")
iv(1, 2, "claimed -> tma issued (offsets)")
iv(2, 3, "tma issued -> front starts (load + queue)")
iv(3, 4, "front: parse")
iv(4, 5, "front: match")
iv(5, 6, "front: exchange + size + publish")
iv(6, 7, "front: stats + meta + hand over")
iv(1, 6, "CLAIM -> PUBLISH (mate 0)")
iv(1, 14, "claim -> publish (mate 1)")
iv(7, 8, "handed over -> scan starts (scan queue)")
iv(8, 9, "scan: look-back")
iv(6, 9, "PUBLISH -> READY")
iv(9, 10, "ready -> back starts (back queue)")
iv(10, 11, "back warp (0,0): emit")
iv(10, 15, "back warp (0,3): emit  [BYP: ready -> copy warp 0 done]")
iv(9, 15, "S ready -> copy warp 0 done (BYP)")
iv(11, 12, "back: wait for the other parts")
iv(12, 13, "back: store + second sync")
iv(10, 13, "BACK total")
iv(0, 13, "SLOT ROUND TRIP")
