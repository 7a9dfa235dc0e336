"""CPU-only tests of the host side: the `barkit` CLI contract, bk_run's error
behaviour without a device, and the multi-rank sharding logic under gloo (world_size 2)."""
import os
import socket
import subprocess
import sys

import pytest

from barkit_b200 import _lib, multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "barkit_b200", "barkit")

FQ = b"@a 1\nACGTACGTACGTAAAA\n+\nIIIIIIIIIIIIIIII\n@b 1\nNNNNACGTACGTAAAA\n+\nIIIIIIIIIIIIIIII\n"


def run_cli(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


@pytest.fixture
def fq(tmp_path):
    p = tmp_path / "in.fq"
    p.write_bytes(FQ)
    return str(p)


def test_cli_exists():
    assert os.access(CLI, os.X_OK), "barkit CLI not built (python -m barkit_b200.build)"


def test_version_and_help():
    assert run_cli("--version").stdout.startswith("barkit ")
    h = run_cli("extract", "--help")
    assert h.returncode == 0
    for flag in ("--fq1", "--fq2", "--out-fq1", "--out-fq2", "--pattern1", "--pattern2", "--gz", "--bgz", "--mgz",
                 "--lz4", "--rc-barcodes", "--skip-trimming", "--max-error", "--threads", "--quiet", "--force"):
        assert flag in h.stdout
    assert run_cli("extract").returncode == 2          # arg_required_else_help (src/lib.rs:29)


@pytest.mark.parametrize("args,needle", [      # clap constraints, src/lib.rs:51-98
    (["-o", "x", "-p", "A"], "--fq1"),
    (["-1", "IN", "-p", "A"], "--out-fq1"),
    (["-1", "IN", "-o", "x"], "--pattern1"),
    (["-1", "IN", "-2", "IN", "-o", "x", "-p", "A"], "--out-fq2"),
    (["-1", "IN", "-o", "x", "-P", "A"], "--fq2"),
    (["-1", "IN", "-o", "x", "-p", "A", "--gz", "--lz4"], "cannot be used together"),
    (["-1", "IN", "-o", "x", "-p", "A", "--bogus"], "unexpected argument"),
    (["-1", "IN", "-o", "x", "-p", "A", "-e", "x"], "--max-error"),
])
def test_usage_errors_exit_2(fq, args, needle):
    r = run_cli("extract", *[fq if a == "IN" else a for a in args])
    assert r.returncode == 2 and needle in r.stderr


def test_max_memory_is_not_global(fq, tmp_path):     # src/lib.rs:9-11
    r = run_cli("extract", "-1", fq, "-o", str(tmp_path / "o"), "-p", "A", "-m", "8")
    assert r.returncode == 2


def test_existing_output_needs_force(fq, tmp_path):  # fastq.rs:219-225, run.rs:108-111
    out = tmp_path / "o.fq"
    out.write_text("x")
    r = run_cli("extract", "-1", fq, "-o", str(out), "-p", "^(?P<UMI>[ATGCN]{4})")
    assert r.returncode == 1
    assert r.stderr.strip() == "I/O error: %s is already existed, use --force to override" % out
    assert r.stdout.splitlines() == ["[1/3] Estimating reads count..."]      # run.rs:97 then exit


def test_bad_pattern_messages(fq, tmp_path):          # error.rs:11-14, run.rs:113-118
    r = run_cli("--quiet", "extract", "-1", fq, "-o", str(tmp_path / "o1"), "-p", "^(?P<XX>[ATGC]{4})")
    assert r.returncode == 1 and r.stderr.strip() == "Provided unexpected barcode capture group XX"
    r = run_cli("--quiet", "extract", "-1", fq, "-o", str(tmp_path / "o2"), "-p", "^atgc")
    assert r.returncode == 1
    assert r.stderr.strip() == "^(ATG.|AT.C|A.GC|.TGC) capture group not found in your pattern"
    r = run_cli("--quiet", "extract", "-1", fq, "-o", str(tmp_path / "o3"), "-p", r"^(?P<UMI>\w+)")
    assert r.returncode == 1 and "unsupported pattern construct" in r.stderr


def test_invalid_shape_message_and_exit_0(fq, tmp_path):   # run.rs:56-58
    import ctypes as C
    a = _lib.RunArgs()
    a.fq1, a.out_fq1, a.out_fq2, a.pattern1 = fq.encode(), str(tmp_path / "o").encode(), str(tmp_path / "O").encode(), b"A"
    err = C.create_string_buffer(256)
    assert _lib.lib.bk_run(C.byref(a), err, 256) == 0


def test_codecs_are_recognised_before_the_device_is_needed(fq, tmp_path):
    """gzip / bgzf / mgzip / lz4 are read and written (tests/test_cli_gpu.py, tests/test_lz4_codec.py); an unknown
    compression selector is an argument error."""
    import ctypes as C
    a = _lib.RunArgs()
    a.fq1, a.out_fq1, a.pattern1, a.compression, a.quiet = fq.encode(), str(tmp_path / "o").encode(), b"^(?P<UMI>[ATGC]{4})", 9, 1
    err = C.create_string_buffer(256)
    assert _lib.lib.bk_run(C.byref(a), err, 256) == _lib.BK_E_ARG and b"compression" in err.value


@pytest.mark.skipif(_lib.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly_not_silently(fq, tmp_path):
    r = run_cli("extract", "-1", fq, "-o", str(tmp_path / "o"), "-p", "^(?P<UMI>[ATGCN]{4})")
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr
    with pytest.raises(_lib.BarkitError) as e:
        _lib.Context(_lib.Program("^(?P<UMI>[ATGCN]{4})", 1))
    assert e.value.status == _lib.BK_E_CUDA


# ---- multi-rank sharding (gloo, world_size 2) ---------------------------------------------------------

def test_round_robin_deal_is_a_partition():
    for nb in (0, 1, 7, 16):
        for world in (1, 2, 4, 8):
            got = sorted(b for r in range(world) for b in multigpu.batches_for_rank(nb, r, world))
            assert got == list(range(nb))
    assert multigpu.batches_of(10, 4) == [(0, 4), (4, 8), (8, 10)]
    firsts = {multigpu.rank_first_index(r, s, 10_000_000) for r in range(8) for s in range(4)}
    assert len(firsts) == 32


_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from barkit_b200 import multigpu
from oracle import bksyn, cport
rank, world, port, outdir = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % port, rank=rank, world_size=world)
cfg = bksyn.CONFIGS["C3"]
pat = "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})"
n, per = 5000, 700
batches = multigpu.batches_of(n, per)
multigpu.barrier(dist)
mine = []
for b in multigpu.batches_for_rank(len(batches), rank, world):
    a, e = batches[b]
    fq1 = cport.generate(cfg, 1, a, e - a, bksyn.DEFAULT_SEED, 1).tobytes()
    fq2 = cport.generate(cfg, 2, a, e - a, bksyn.DEFAULT_SEED, 1).tobytes()
    r = cport.extract(fq1, fq2, pat, None, 1)       # the oracle stands in for the GPU on this CPU box
    mine.append((b, r["out1"], r["out2"]))
multigpu.barrier(dist)
t = multigpu.max_over_ranks(10.0 + rank, dist)     # ranks report different times; everyone must see the max
assert t == 10.0 + world - 1, t
import pickle
pickle.dump(mine, open(os.path.join(outdir, "rank%d.pkl" % rank), "wb"))
dist.destroy_process_group()
'''


def test_two_rank_run_reassembles_to_single_rank_bytes(tmp_path):
    import pickle
    from oracle import bksyn, cport
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(port), str(tmp_path)])
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    parts = []
    for r in range(2):
        parts += pickle.load(open(tmp_path / ("rank%d.pkl" % r), "rb"))
    out1 = multigpu.reassemble([(b, o1) for b, o1, _ in parts])
    out2 = multigpu.reassemble([(b, o2) for b, _, o2 in parts])
    cfg = bksyn.CONFIGS["C3"]
    fq1 = cport.generate(cfg, 1, 0, 5000, bksyn.DEFAULT_SEED, 2).tobytes()
    fq2 = cport.generate(cfg, 2, 0, 5000, bksyn.DEFAULT_SEED, 2).tobytes()
    ref = cport.extract(fq1, fq2, "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", None, 1)
    assert out1 == ref["out1"] and out2 == ref["out2"]
