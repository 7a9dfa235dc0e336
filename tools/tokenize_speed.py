"""Throughput of bk_tokenize_device on device-resident bksyn-v1 text, wall-timed around
the call (it ends with a small device -> host read of the counts), against the host tokenizer on one core.
  python tools/tokenize_speed.py [records]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
from tests.util import lib_syn_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
rl = 329
ctx = _lib.Context(_lib.Program("^(?P<UMI>[ATGCN]{12})", 1))
d_text = ctx.dev_alloc(P * rl + 64)
d_ref = ctx.dev_alloc((P + 1) * 4)
ctx.generate_device(lib_syn_config("C1"), 1, 20240902, 0, P, d_text, d_ref)
d_off = ctx.dev_alloc((P + 2) * 4)
n, consumed, maxrec = C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
best = 1e9
for it in range(6):
    t0 = time.perf_counter()
    rc = _lib.lib.bk_tokenize_device(ctx.handle, d_text, P * rl, 1, d_off, P + 1, C.byref(n), C.byref(consumed), C.byref(maxrec))
    dt = time.perf_counter() - t0
    assert rc == 0 and n.value == P and consumed.value == P * rl and maxrec.value == rl, (rc, n.value, consumed.value, maxrec.value)
    if it:
        best = min(best, dt)
same = (ctx.d2h(d_off, (P + 1) * 4, np.uint32) == ctx.d2h(d_ref, (P + 1) * 4, np.uint32)).all()
print("bk_tokenize_device: %d records, %.1f MB: %.3f ms = %.0f GB/s of text, %.1f G records/s; offsets identical to the generator's: %s"
      % (P, P * rl / 1e6, best * 1e3, P * rl / best / 1e9, P / best / 1e9, bool(same)))
Q = min(P, 2000000)
text = ctx.d2h(d_text, Q * rl)
off = np.zeros(Q + 2, np.uint32)
t0 = time.perf_counter()
rc = _lib.lib.bk_tokenize(text.ctypes.data, Q * rl, 1, off.ctypes.data, Q + 1, C.byref(n), C.byref(consumed), C.byref(maxrec))
dt = time.perf_counter() - t0
print("bk_tokenize (host, 1 core): %d records: %.1f ms = %.2f GB/s" % (Q, dt * 1e3, Q * rl / dt / 1e9))
