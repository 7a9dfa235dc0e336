"""Do the host cores' copies and the GPU's DMA share one ceiling?  Plain copies on all host cores (bk_host_mem_probe)
alone, the host link alone (bk_bus_probe), then both at once.   python tools/mem_vs_bus.py [threads]"""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
ctx = _lib.Context(_lib.Program("^(?P<UMI>[ATGCN]{12})", 0), None, paired=False)
def mem(iters=2):
    g = C.c_double(0.0)
    assert _lib.lib.bk_host_mem_probe(64 << 20, T, iters, C.byref(g)) == 0
    return g.value
print("cores alone   : %.1f GB/s (read + written, %d threads)" % (mem(3), T), flush=True)
print("link alone    : h2d %.1f  d2h %.1f  duplex %.1f GB/s" % ctx.bus_probe(256 << 20, 8), flush=True)
out = {}
th = threading.Thread(target=lambda: out.setdefault("bus", ctx.bus_probe(256 << 20, 60)))
th.start()
vals = []
while th.is_alive():
    vals.append(mem(1))
th.join()
print("both at once  : link h2d %.1f  d2h %.1f  duplex %.1f GB/s; cores %s GB/s (one figure per probe while the link ran)"
      % (out["bus"] + (" ".join("%.0f" % v for v in vals),)), flush=True)
ctx.close()
