"""In-tree build of the CUDA library and the `barkit` CLI for sm_100a (nvcc cross-compiles
without a GPU).  Artefacts: barkit_b200/libbarkit_b200.so, barkit_b200/barkit."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbarkit_b200.so")
CLI = os.path.join(HERE, "barkit")

LIB_SOURCES = ["bk_ctx.cu", "bk_pattern.cc", "bk_fastq.cc", "bk_run.cc"]
CLI_SOURCES = ["barkit_cli.cc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-pthread"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = list(sources)
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        deps += [os.path.join(d, f) for f in os.listdir(d)]
    return any(os.path.getmtime(s) > t for s in deps if os.path.isfile(s))


def build(force=False, verbose=False):
    nvcc = _nvcc()
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force or _stale(LIB, srcs):
        cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB] + srcs + ["-lpthread", "-lz"]
        if os.environ.get("BK_PROFILE_PHASES"):
            cmd.insert(1, "-DBK_PROFILE_PHASES")
        for d in os.environ.get("BK_DEFINES", "").split():
            cmd.insert(1, "-D" + d)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    cli_srcs = [os.path.join(CSRC, s) for s in CLI_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if cli_srcs and (force or _stale(CLI, cli_srcs + [LIB])):
        cmd = [nvcc] + NVCC_FLAGS + ["-o", CLI] + cli_srcs + [
            "-L" + HERE, "-lbarkit_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
