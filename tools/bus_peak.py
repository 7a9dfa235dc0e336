"""The box's host-link ceiling: pinned host <-> device copies on 1, 2, 4, ... GPUs at once (one context and one
host thread per GPU, bk_bus_probe: H2D alone, D2H alone, both directions together).  This is the denominator of
the end-to-end numbers (bench.py's e2e.bus_peak_gbs is the same probe run by every rank at once).
  python tools/bus_peak.py [MiB per copy] [iterations]"""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ndev = _lib.device_count()
if ndev < 1:
    raise SystemExit("no CUDA device")
prog = _lib.Program("^(?P<UMI>[ATGCN]{12})", 1)
ctxs = [_lib.Context(prog, device=d) for d in range(ndev)]
print("GPUs visible: %d, host threads: %d, %d MiB per copy, %d copies per direction" % (ndev, os.cpu_count(), mib, iters))
print("%5s %12s %12s %14s   per-GPU duplex (GB/s)" % ("GPUs", "H2D GB/s", "D2H GB/s", "duplex GB/s"))
g = 1
while g <= ndev:
    res = [None] * g
    bar = threading.Barrier(g)

    def work(i):
        bar.wait()
        res[i] = ctxs[i].bus_probe(mib << 20, iters)

    th = [threading.Thread(target=work, args=(i,)) for i in range(g)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    print("%5d %12.1f %12.1f %14.1f   %s" % (g, sum(r[0] for r in res), sum(r[1] for r in res), sum(r[2] for r in res),
                                           " ".join("%.0f" % r[2] for r in res)), flush=True)
    g *= 2
for c in ctxs:
    c.close()
