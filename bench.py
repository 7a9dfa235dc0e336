#!/usr/bin/env python
"""Benchmark of the `barkit extract` hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

Workload (BASELINE.json configs[2], "C3"): 10x-v3-style R1 = CB16 + fuzzy linker + UMI12,
150 bp paired reads, bksyn-v1 synthetic text generated on the device.  One step = one batch of
`--pairs` read pairs; 10 steps of 10 M pairs = the 100 M pairs the config names.

  value     read pairs/s, device-timed (CUDA events around every launch, inputs and outputs in HBM)
  e2e       the same call through host buffers: pinned FASTQ text -> H2D -> kernel -> D2H
  roofline  algorithmic bytes (input FASTQ text + emitted FASTQ text) / kernel time vs measured HBM peak
  cpu_baseline  oracle/cpu_extract.c (a restatement of the reference; the reference is Rust and
            cannot be built here) on all host cores, bounded sample
Multi-GPU: one process per GPU (torchrun), every rank runs the same number of independent batches
(weak scaling, no collective on the data path); torch.distributed is used only for the barrier and
the max-over-ranks of the device time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATTERN1 = "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})"
MAX_ERROR = 1
READ_LEN = 150
SEED = 20240902
METRIC = "read pairs/sec extracted (device-timed)"
UNIT = "pairs/s"
WORKLOAD = ("C3: 10x-v3-style R1 CB16+atgccat(1 mismatch)+UMI12, 150bp paired reads, bksyn-v1 synthetic, "
            "-p '%s' -e %d" % (PATTERN1, MAX_ERROR))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 10 ms from before the warm-up on; only the samples
    whose time stamp falls inside the timed region count (widened to the warm-up steps -- the same load -- when
    the region is too short to catch two)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None
        self.t_start = self.t_begin = self.t_end = None

    def start(self):
        import datetime
        self.t_start = datetime.datetime.now()
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "10"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.3)   # nvidia-smi needs a moment before its first sample
        except Exception:
            self.proc = None

    def begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def end(self):
        import datetime
        self.t_end = datetime.datetime.now()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "timed region"}
        if not self.proc:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        lo, hi = self.t_begin or self.t_start, self.t_end or datetime.datetime.now()
        sel = [r for r in rows if lo <= r[0] <= hi]
        if len(sel) < 2:   # a 30 ms region at a 10 ms period: take the warm-up steps (same kernel, same inputs) too
            sel = [r for r in rows if self.t_start + datetime.timedelta(seconds=0.3) <= r[0] <= hi]
            out["window"] = "warm-up + timed region"
        sm = sorted(r[1] for r in sel)
        reasons = set()
        for r in sel:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in sel), reasons=sorted(reasons), samples=len(sm))
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


_CPU_OUT = {}


def cpu_port_rate(text1, off1, text2, off2, n_pairs, threads):
    """pairs/s of the C restatement on `threads` host threads over the first n_pairs pairs.
    Output buffers are allocated and touched once, outside the clock (the reference's writer owns
    a long-lived buffer too), so the figure is the algorithm, not first-touch page faults."""
    import numpy as np
    from oracle import cport
    a1, a2 = text1[:int(off1[n_pairs])], text2[:int(off2[n_pairs])]
    o1, o2 = off1[:n_pairs + 1], off2[:n_pairs + 1]
    need = a1.size + 80 * n_pairs + 4096
    if _CPU_OUT.get("size", 0) < need:
        _CPU_OUT["bufs"] = (np.zeros(need, np.uint8), np.zeros(need, np.uint8))
        _CPU_OUT["size"] = need
        cport.extract(a1, a2, PATTERN1, None, MAX_ERROR, threads=threads, off1=o1, off2=o2,
                      out_bufs=_CPU_OUT["bufs"])     # untimed: faults the scratch arena in
    t0 = time.perf_counter()
    r = cport.extract(a1, a2, PATTERN1, None, MAX_ERROR, threads=threads, off1=o1, off2=o2,
                      out_bufs=_CPU_OUT["bufs"])
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, r


def run_reference(args, rank):
    """--impl reference: the reference algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import numpy as np
    from oracle import bksyn, cport
    threads = host_threads()
    cfg = bksyn.CONFIGS["C3"]
    rl = bksyn.record_len(READ_LEN)
    # size one step for ~6 s of CPU work after a short probe
    probe = 200_000
    t1 = cport.generate(cfg, 1, 0, probe, SEED, threads)
    t2 = cport.generate(cfg, 2, 0, probe, SEED, threads)
    off = (np.arange(probe + 1, dtype=np.uint64) * rl).astype(np.uint32)
    rate, _, _ = cpu_port_rate(t1, off, t2, off, probe, threads)
    n = int(min(max(rate * 6.0, 200_000), 6_000_000))
    t1 = cport.generate(cfg, 1, 0, n, SEED, threads)
    t2 = cport.generate(cfg, 2, 0, n, SEED, threads)
    off = (np.arange(n + 1, dtype=np.uint64) * rl).astype(np.uint32)
    for _ in range(args.warmup):
        cpu_port_rate(t1, off, t2, off, n, threads)
    total = 0.0
    for _ in range(args.steps):
        _, dt, r = cpu_port_rate(t1, off, t2, off, n, threads)
        total += dt
    value = args.steps * n / total
    sample = "%d pairs per step (bounded sample of the C3 workload), pre-tokenised, in host memory" % n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n, "read_len": READ_LEN},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle/cpu_extract.c: C restatement of the reference (Rust; no toolchain in "
                                 "this image) with ideal chunked parallelism = upper bound of its structure"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per step (device-resident)")
    ap.add_argument("--e2e-pairs", type=int, default=2_000_000, help="read pairs per e2e step (host buffers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29513"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    from barkit_b200 import _lib
    from oracle import bksyn  # only for the record-layout constants of the synthetic generator

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from barkit_b200 import multigpu
    dev = "cuda:%d" % local_rank if dist is not None else None

    def barrier():
        multigpu.barrier(dist, dev)

    def max_over_ranks(x):
        return multigpu.max_over_ranks(x, dist, dev)

    if _lib.device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device %d; the extract path has no CPU fallback" % local_rank)
    prog = _lib.Program(PATTERN1, MAX_ERROR)
    # one process per GPU: the host threads of the library (output of the pattern-less mate) are shared out
    # among the ranks of this box instead of every rank starting one per core
    os.environ.setdefault("BK_HOST_THREADS", str(max(2, (os.cpu_count() or 16) // max(world, 1))))
    ctx = _lib.Context(prog, None, paired=True, device=local_rank)
    cfg = _lib.syn_config(READ_LEN, plant1=(b"ATGCCAT", 16, 0))
    rl = bksyn.record_len(READ_LEN)
    P = args.pairs
    nbatch = max(1, min(4, args.steps + args.warmup))

    def make_batch(first_idx, n):
        mates = []
        for mate in (1, 2):
            d_text = ctx.dev_alloc(n * rl + 64)
            d_off = ctx.dev_alloc((n + 1) * 4)
            ctx.generate_device(cfg, mate, SEED, first_idx, n, d_text, d_off)
            m = _lib.MateIn()
            m.text, m.rec_off, m.text_len, m.max_rec_bytes = d_text, d_off, n * rl, rl
            mates.append(m)
        return mates

    # distinct batches per rank and per step slot: rank r owns records [r*2^36 + b*P, ...)
    batches = [make_batch(multigpu.rank_first_index(rank, b, P), P) for b in range(nbatch)]
    cap1 = ctx.out_bound(1, P, P * rl)
    cap2 = ctx.out_bound(2, P, P * rl)
    d_o1, d_o2 = ctx.dev_alloc(cap1), ctx.dev_alloc(cap2)

    def step(i):
        m = batches[i % nbatch]
        return ctx.extract_device(P, m[0], m[1], d_o1, cap1, d_o2, cap2)

    # set-up, untimed: with BK_KERNEL=auto a context whose patterns qualify for both extract kernels times them
    # on its first large batches (two launches each) and keeps the faster; let it settle before the warm-up
    # (by default every launch uses extract_kernel and this loop does nothing)
    for i in range(8):
        if ctx.kernel_variant() != 0:
            break
        step(i)
    variant = ctx.kernel_variant()
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.begin()
    t_wall0 = time.perf_counter()
    dev_ms, launches, alg_bytes, kept = 0.0, 0, 0, 0
    for i in range(args.steps):
        r = step(args.warmup + i)
        dev_ms += r.kernel_ms
        launches += r.launches
        alg_bytes += 2 * P * rl + r.out1_len + r.out2_len
        kept += r.n_out
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    sampler.end()
    clocks = sampler.stop()
    dev_ms_max = max_over_ranks(dev_ms)
    value = world * args.steps * P / (dev_ms_max / 1e3)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (dev_ms / 1e3) / 1e9

    # ---- e2e through host buffers (pinned), double-buffered over the ctx's two slots ---------------
    e2e = None
    if not args.no_e2e:
        Q = min(args.e2e_pairs, P)
        nb = Q * rl
        hbuf = {}
        for name, size in (("t1", nb + 64), ("t2", nb + 64), ("o1a", ctx.out_bound(1, Q, nb)),
                           ("o2a", ctx.out_bound(2, Q, nb)), ("o1b", ctx.out_bound(1, Q, nb)),
                           ("o2b", ctx.out_bound(2, Q, nb)), ("off", (Q + 1) * 4)):
            p = _lib.lib.bk_host_alloc(size)
            if not p:
                raise SystemExit("pinned allocation failed")
            hbuf[name] = (p, size)
        m = batches[0]
        _lib.lib.bk_memcpy_d2h(ctx.handle, hbuf["t1"][0], m[0].text, nb)
        _lib.lib.bk_memcpy_d2h(ctx.handle, hbuf["t2"][0], m[1].text, nb)
        _lib.lib.bk_memcpy_d2h(ctx.handle, hbuf["off"][0], m[0].rec_off, (Q + 1) * 4)
        h1, h2 = _lib.MateIn(), _lib.MateIn()
        for h, t in ((h1, "t1"), (h2, "t2")):
            h.text, h.rec_off, h.text_len, h.max_rec_bytes = hbuf[t][0], hbuf["off"][0], nb, rl
        outs = [("o1a", "o2a"), ("o1b", "o2b")]
        res = _lib.Result()

        def submit(slot):
            rc = _lib.lib.bk_submit(ctx.handle, slot, Q, C.byref(h1), C.byref(h2))
            if rc != 0:
                raise SystemExit("bk_submit failed: %s" % _lib.lib.bk_last_error(ctx.handle).decode())

        def wait(slot):
            a, b = outs[slot]
            rc = _lib.lib.bk_wait(ctx.handle, slot, hbuf[a][0], hbuf[a][1], hbuf[b][0], hbuf[b][1], C.byref(res))
            if rc != 0:
                raise SystemExit("bk_wait failed: %s" % _lib.lib.bk_last_error(ctx.handle).decode())
            return (res.h2d_bytes, res.d2h_bytes)

        def e2e_run(nsteps):
            d2h = (0, 0)
            submit(0)
            for s in range(nsteps):
                if s + 1 < nsteps:
                    submit((s + 1) & 1)
                d2h = wait(s & 1)
            return d2h

        e2e_run(max(2, min(args.warmup, 3)))
        barrier()
        t0 = time.perf_counter()
        d2h = e2e_run(args.steps)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * args.steps * Q / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(d2h[0]), "d2h_bytes_per_step": int(d2h[1]),
               "pairs_per_step": Q, "ms_per_step": 1e3 * e2e_s / args.steps,
               "host_threads": int(os.environ.get("BK_HOST_THREADS", "0")),
               "note": "bk_submit/bk_wait with pinned host buffers, 2 slots in flight; h2d/d2h bytes as the "
                       "library counts them (text + record offsets in, FASTQ text + result + match table out). "
                       "R2 has no pattern in this config: it stays on the host and its output (the kept "
                       "pairs' records) is assembled there from the input buffer, inside the timed region"}

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        probe = min(200_000, P)
        m = batches[0]
        t1 = ctx.d2h(m[0].text, min(P, 6_000_000) * rl)
        t2 = ctx.d2h(m[1].text, min(P, 6_000_000) * rl)
        off = (np.arange(min(P, 6_000_000) + 1, dtype=np.uint64) * rl).astype(np.uint32)
        rate, _, _ = cpu_port_rate(t1, off, t2, off, probe, threads)
        n = int(min(max(rate * 12.0, probe), min(P, 6_000_000)))
        # ~12 s of CPU work: as many passes over the n-pair sample as that takes (capped)
        _, dt1, _ = cpu_port_rate(t1, off, t2, off, n, threads)
        reps = int(max(1, min(200, round(12.0 / max(dt1, 1e-3)))))
        total = 0.0
        for _ in range(reps):
            _, dt, r = cpu_port_rate(t1, off, t2, off, n, threads)
            total += dt
        rate = reps * n / total
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d passes over the first %d pairs of step 0's batch (%.1f s), pre-tokenised, "
                         "host memory -> host memory" % (reps, n, total),
               "note": "oracle/cpu_extract.c, C restatement of the reference (Rust toolchain absent); "
                       "ideal chunked parallelism, so an upper bound of the reference's structure"}

    if rank == 0:
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            # dram bytes per launch from the committed ncu --set full capture, scaled to this batch size
            traffic = tj["dram_bytes_per_pair"] * P
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": P, "read_len": READ_LEN,
                       "kept_fraction": kept / float(args.steps * P),
                       "l2": "each step reads %.1f GB of fresh input (>> 126 MB L2); %d distinct resident "
                             "batches cycled; no explicit flush" % (2 * P * rl / 1e9, nbatch),
                       "wall_ms_per_step": wall_ms / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_step": alg_bytes / args.steps,
                         "kernel": ("bk::v5::extract_kernel5<PAIRED=true>" if variant == 5 else
                                    "bk::extract_kernel<PAIRED=true, GENERAL=false>") + " (1 launch per step)"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
