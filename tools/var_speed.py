import sys, os
sys.path.insert(0, "/root/repo")
from barkit_b200 import _lib
from tests.util import lib_syn_config
P = 4000000; rl = 329
for pat, k in (("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", 1), ("^(?P<CB>[ATGCN]{14,16})atgccat(?P<UMI>[ATGCN]{12})", 1),
               ("^(?P<CB>[ATGCN]{14,16})atgccat(?P<UMI>[ATGCN]{10,12})", 1), ("(?P<CB>[ATGCN]{14,16})atgccat", 1)):
    prog = _lib.Program(pat, k)
    ctx = _lib.Context(prog, None, paired=True)
    ms = []
    for mate in (1, 2):
        dt = ctx.dev_alloc(P * rl + 64); do = ctx.dev_alloc((P + 1) * 4)
        ctx.generate_device(lib_syn_config("C3"), mate, 20240902, 0, P, dt, do)
        m = _lib.MateIn(); m.text, m.rec_off, m.text_len, m.max_rec_bytes = dt, do, P * rl, rl; ms.append(m)
    cap = P * (rl + 120)
    o1 = ctx.dev_alloc(cap); o2 = ctx.dev_alloc(cap)
    for i in range(6):
        r = ctx.extract_device(P, ms[0], ms[1], o1, cap, o2, cap)
    print("%-60s variants %2d: %7.3f ms %7.1f M pairs/s kept %.3f" % (pat, prog.nvariants(), r.kernel_ms, P / r.kernel_ms / 1e3, r.n_out / P), flush=True)
    ctx.close()
