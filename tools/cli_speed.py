"""Wall time of the `barkit extract` command line, file -> file on tmpfs (end to end,
reported beside the device-timed number).  Writes a bksyn-v1 C3 pair of FASTQ files to /dev/shm first.
  python tools/cli_speed.py [pairs]"""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
from tests.util import CONFIG_PATTERNS, lib_syn_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
rl = 329
p1, p2, _ = CONFIG_PATTERNS["C3"]
d = "/dev/shm/bk_cli_speed"
os.makedirs(d, exist_ok=True)
ctx = _lib.Context(_lib.Program(p1, 1), None, paired=True)
for mate in (1, 2):
    dt = ctx.dev_alloc(P * rl + 64); do = ctx.dev_alloc((P + 1) * 4)
    ctx.generate_device(lib_syn_config("C3"), mate, 20240902, 0, P, dt, do)
    ctx.d2h(dt, P * rl).tofile(os.path.join(d, "r%d.fq" % mate))
    ctx.dev_free(dt); ctx.dev_free(do)
ctx.close()
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "barkit_b200", "barkit")
for run in range(3):
    for f in ("o1.fq", "o2.fq"):
        try: os.remove(os.path.join(d, f))
        except FileNotFoundError: pass
    t0 = time.perf_counter()
    r = subprocess.run([cli, "--quiet", "extract", "-1", d + "/r1.fq", "-2", d + "/r2.fq", "-o", d + "/o1.fq", "-O", d + "/o2.fq",
                        "-p", p1, "-e", "1"] + (["--verbose"] if os.environ.get("BK_DEBUG") else []) + os.environ.get("BK_CLI_ARGS", "").split(), capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr
    if os.environ.get("BK_DEBUG"):
        print(r.stderr.strip())
    osz = os.path.getsize(d + "/o1.fq") + os.path.getsize(d + "/o2.fq")
    if run == 0:
        import zlib
        crc = [zlib.crc32(np.fromfile(d + "/" + f, np.uint8).tobytes()) for f in ("o1.fq", "o2.fq")]
        print("output crc32 (identical for any number of GPUs):", crc, "GPUs visible:", _lib.device_count())
    print("barkit extract: %d pairs, %.2f GB in, %.2f GB out: %.3f s = %.1f M pairs/s, %.2f GB/s in+out"
          % (P, 2 * P * rl / 1e9, osz / 1e9, dt, P / dt / 1e6, (2 * P * rl + osz) / dt / 1e9), flush=True)
for f in os.listdir(d):
    os.remove(os.path.join(d, f))
