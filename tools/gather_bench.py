"""Host-side assembly of the pattern-less mate (bk::gather_kept, bk_fastq.cc) timed on this machine's cores, no GPU:
   python tools/gather_bench.py [records] [threads]
Prints GB/s of read + written bytes next to a plain multi-threaded copy of the same size (bk_host_mem_probe)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bksyn
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "barkit_b200", "libbarkit_b200.so"))
g = lib._ZN2bk11gather_keptEPKhPKjjPKiPhmi
g.restype = C.c_uint64
g.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
_, fq2 = bksyn.generate_pair(bksyn.CONFIGS["C3"], 0, 2000)
rec = np.frombuffer(fq2, np.uint8)
rl = len(fq2) // 2000
reps = (N + 1999) // 2000
text = np.tile(rec, reps)[: N * rl].copy()
off = (np.arange(N + 1, dtype=np.uint64) * rl).astype(np.uint32)
rng = np.random.default_rng(1)
keep = np.where(rng.random(N) < 0.8985, 0, -1).astype(np.int32)
out = np.zeros(N * rl + 64, np.uint8)
best = 1e9
for it in range(7):
    t0 = time.perf_counter()
    n = g(text.ctypes.data, off.ctypes.data, N, keep.ctypes.data, out.ctypes.data, out.size, T)
    dt = time.perf_counter() - t0
    if it: best = min(best, dt)
kept = int((keep >= 0).sum())
assert n == kept * rl, (n, kept * rl)
# spot check: the first kept records are where they should be
i = int(np.argmax(keep >= 0))
assert bytes(out[:rl]) == bytes(text[i * rl:(i + 1) * rl])
moved = kept * rl * 2
print("gather_kept: %d records (%d kept), %d threads: %.2f ms = %.1f GB/s read + written" % (N, kept, T, best * 1e3, moved / best / 1e9))
gb = C.c_double(0)
lib.bk_host_mem_probe.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double)]
lib.bk_host_mem_probe(C.c_uint64(max(4096, N * rl // T)), T, 3, C.byref(gb))
print("plain copy : %.1f GB/s read + written (%d threads)" % (gb.value, T))
