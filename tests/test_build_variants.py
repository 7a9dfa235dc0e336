"""The documented build parameters ("Build parameters") still compile for sm_100a.

No GPU needed: nvcc cross-compiles bk_ctx.cu to a cubin in a scratch directory (the in-tree library is not touched),
all variants at once.  What they do on the device is measured with tools/exp_defs.sh on a B200."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "barkit_b200", "csrc")

VARIANTS = [
    ["BK_SE_BACKS=7", "BK_EMIT_CP=0"],          # split seq / qual lines through the stream writer
    ["BK_SE_BACKS=5", "BK_SE_FRONTS=4"],
    ["BK_EMIT_CP=1", "BK_CP_ROT=1", "BK_MATCH_UNROLL=2"],
    ["BK_EXP_NOLB", "BK_EXP_NOEMIT", "BK_PROFILE_PHASES"],   # timing-only probes + the event trace
]


def test_build_parameters_compile():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for i, defs in enumerate(VARIANTS):
            cmd = [nvcc] + ["-D" + d for d in defs] + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                                                        "-cubin", "-o", os.path.join(tmp, "v%d.cubin" % i),
                                                        os.path.join(CSRC, "bk_ctx.cu")]
            procs.append((defs, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for defs, p in procs:
            out, _ = p.communicate(timeout=600)
            assert p.returncode == 0, "%s does not compile:\n%s" % (" ".join(defs), out[-2000:])
