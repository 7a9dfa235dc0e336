"""Device-timed throughput of the extract kernel on every BASELINE.json config (bksyn-v1, generated on the device).
  python tools/config_speed.py [pairs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
from tests.util import CONFIG_PATTERNS, lib_syn_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
rl = 329
ONLY = os.environ.get("BK_ONLY")   # e.g. BK_ONLY=C4: one config (k = 1), for profiling
for name, k in (("C1", 1), ("C2", 1), ("C3", 1), ("C4", 0), ("C4", 1), ("C4", 2)):
    if ONLY and (name != ONLY or k != 1):
        continue
    p1, p2, paired = CONFIG_PATTERNS[name]
    prog1 = _lib.Program(p1, k) if p1 else None
    prog2 = _lib.Program(p2, k) if p2 else None
    ctx = _lib.Context(prog1, prog2, paired=paired, extra_flags=int(os.environ.get("BK_CTX_FLAGS", "0")))  # e.g. 256: BK_FLAG_SPLIT_TWO_PATTERNS
    cfg = lib_syn_config(name)
    ms = []
    for mate in ((1, 2) if paired else (1,)):
        dt = ctx.dev_alloc(P * rl + 64); do = ctx.dev_alloc((P + 1) * 4)
        ctx.generate_device(cfg, mate, 20240902, 0, P, dt, do)
        m = _lib.MateIn(); m.text, m.rec_off, m.text_len, m.max_rec_bytes = dt, do, P * rl, rl; ms.append(m)
    cap = P * (rl + 120)
    o1 = ctx.dev_alloc(cap); o2 = ctx.dev_alloc(cap) if paired else None
    for i in range(5):
        r = ctx.extract_device(P, ms[0], ms[1] if paired else None, o1, cap, o2, cap if paired else 0)
    nbytes = (len(ms) * P * rl + r.out1_len + r.out2_len)
    print("%s k=%d: %8.3f ms  %7.1f M units/s  kept %.3f  %6.0f GB/s" %
          (name, k, r.kernel_ms, P / r.kernel_ms / 1e3, r.n_out / P, nbytes / r.kernel_ms / 1e6), flush=True)
    ctx.close()
