"""Device-timed speed of the single-end kernel on ONE mate of a paired config (what a two-kernel form -- single-end
extract of the patterned mate + a compaction of the other -- would spend on its first kernel).
  python tools/se_speed.py [reads]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
from tests.util import CONFIG_PATTERNS, lib_syn_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
rl = 329
for name, mate in (("C3", 1), ("C2", 2), ("C1", 1)):
    p1, p2, _ = CONFIG_PATTERNS[name]
    pat = p1 if mate == 1 else p2
    ctx = _lib.Context(_lib.Program(pat, 1), None, paired=False)
    dt = ctx.dev_alloc(P * rl + 64); do = ctx.dev_alloc((P + 1) * 4); dm = ctx.dev_alloc((P + 1) * 4)
    ctx.generate_device(lib_syn_config(name), mate, 20240902, 0, P, dt, do)
    m = _lib.MateIn(); m.text, m.rec_off, m.text_len, m.max_rec_bytes = dt, do, P * rl, rl
    cap = P * (rl + 120); o1 = ctx.dev_alloc(cap)
    for with_match in (0, 1):
        for i in range(4):
            r = ctx.extract_device(P, m, None, o1, cap, 0, 0, d_match1=dm if with_match else 0)
        nbytes = P * rl + r.out1_len
        print("%s mate %d single-end%s: %7.3f ms  %7.1f M reads/s  kept %.3f  %6.0f GB/s" %
              (name, mate, " + match table" if with_match else "", r.kernel_ms, P / r.kernel_ms / 1e3, r.n_out / P, nbytes / r.kernel_ms / 1e6), flush=True)
    ctx.close()
