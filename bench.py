#!/usr/bin/env python
"""Benchmark of the `barkit extract` hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

Workload (BASELINE.json configs[2], "C3"): 10x-v3-style R1 = CB16 + fuzzy linker + UMI12,
150 bp paired reads, bksyn-v1 synthetic text generated on the device.  One step = one batch of
`--pairs` read pairs; 10 steps of 10 M pairs = the 100 M pairs the config names.

  value     read pairs/s, device-timed (CUDA events around every launch, inputs and outputs in HBM)
  e2e       the same call through host buffers: pinned FASTQ text -> H2D -> kernel -> D2H, in the library's
            default mode and with the pattern-less mate pinned to the host / to the device; next to it the box's
            measured host-link ceiling (bus_peak_gbs: concurrent pinned H2D + D2H on all ranks' GPUs)
  roofline  algorithmic bytes (input FASTQ text + emitted FASTQ text) / kernel time vs measured HBM peak
  configs   every other BASELINE config (C1, C2, C4 at -e 0/1/2) and the device tokenizer, device-timed the
            same way (N = 1 only)
  file_e2e  the product's own file -> file path: the `barkit` CLI on tmpfs with --gpus N (bk_run's dispatcher,
            per-GPU workers and ordered writer), its output CRC against the --gpus 1 run, and the CPU port
            doing the same files beside it
  cpu_baseline  oracle/cpu_extract.c (a restatement of the reference; the reference is Rust and
            cannot be built here -- `cargo` is probed at run time) on all host cores, bounded sample
Multi-GPU: one process per GPU (torchrun), every rank runs the same number of independent batches
(weak scaling, no collective on the data path); torch.distributed is used only for the barrier and
the max-over-ranks of the device time.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATTERN1 = "^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})"
MAX_ERROR = 1
READ_LEN = 150
RL = 23 + 2 * READ_LEN + 6          # bksyn-v1 record bytes (oracle/bksyn.py)
SEED = 20240902
METRIC = "read pairs/sec extracted (device-timed)"
UNIT = "pairs/s"
WORKLOAD = ("C3: 10x-v3-style R1 CB16+atgccat(1 mismatch)+UMI12, 150bp paired reads, bksyn-v1 synthetic, "
            "-p '%s' -e %d" % (PATTERN1, MAX_ERROR))
CLI = os.path.join(ROOT, "barkit_b200", "barkit")

# the other BASELINE.json configs: name -> (pattern1, pattern2, paired, plant1, plant2)  
OTHER_CONFIGS = {
    "C1": ("^(?P<UMI>[ATGCN]{12})", None, False, None, None),
    "C2": (None, "^(?P<CB>[ATGCN]{16})atgccat", True, None, (b"ATGCCAT", 16, 0)),
    "C4": ("^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", "(?P<CB>[ATGCN]{16})atgccat", True,
           (b"GTCAGTAC", 8, 0), (b"ATGCCAT", 16, 40)),
    # SURVEY.md section 8f-2 (not BASELINE configs): variable repetition on C3's data, pattern on R1 only
    "F2_unanchored_3_variants": ("(?P<CB>[ATGCN]{14,16})atgccat", None, True, (b"ATGCCAT", 16, 0), None),
    "F2_unbounded_greedy": ("^(?P<CB>[ATGCN]{14,})atgccat(?P<UMI>[ATGCN]{12})", None, True, (b"ATGCCAT", 16, 0), None),
    # the general form (a repeated lowercase run: a backtracking program run by match_kernel<VM>)
    "F2_general_repeated_run": ("^(?P<CB>[ATGCN]{16})(atgccat)+(?P<UMI>[ATGCN]{12})", None, True, (b"ATGCCAT", 16, 0), None),
}
F2_KERNELS = ("bk::match_kernel on the pattern's mate + bk::extract_kernel<PAIRED=false, VAR=true, TABLE=true> on it + "
              "bk::compact_kernel on the other")


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


# ---- the reference on the host cores ---------------------------------------------------------------------------

def reference_binary():
    """prefer the real Rust binary.  -> (path or None, why).  Builds /root/reference with
    cargo (offline) into oracle/_ref/ when both exist; on the GPU box neither does and the C port stands in."""
    exe = os.path.join(ROOT, "oracle", "_ref", "release", "barkit")
    if os.path.exists(exe):
        return exe, "oracle/_ref/release/barkit (built from /root/reference)"
    cargo = shutil.which("cargo")
    if not cargo:
        return None, "cargo not on PATH"
    src = "/root/reference"
    if not os.path.isdir(src):
        return None, "cargo present but the reference sources are not on this box"
    try:
        r = subprocess.run([cargo, "build", "--release", "--locked", "--offline", "--manifest-path",
                            os.path.join(src, "Cargo.toml"), "--target-dir", os.path.join(ROOT, "oracle", "_ref")],
                           capture_output=True, text=True, timeout=900)
        if r.returncode == 0 and os.path.exists(exe):
            return exe, "oracle/_ref/release/barkit (built from /root/reference)"
        return None, "cargo build failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200]
    except Exception as e:   # noqa: BLE001
        return None, "cargo build failed: %s" % e


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
    """--impl reference: the reference algorithm on the host cores (rank 0 only), on the same config and the same
    pairs per step as the CUDA arm."""
    if rank != 0:
        return
    import numpy as np
    from oracle import bksyn, cport
    threads = host_threads()
    cfg = bksyn.CONFIGS["C3"]
    exe, why = reference_binary()
    n = args.pairs
    t1 = cport.generate(cfg, 1, 0, n, SEED, threads)
    t2 = cport.generate(cfg, 2, 0, n, SEED, threads)
    off = (np.arange(n + 1, dtype=np.uint64) * RL).astype(np.uint32)
    kind = "port"
    if exe:   # the real thing: file -> file on tmpfs, all cores through RAYON_NUM_THREADS
        d = tempfile.mkdtemp(prefix="bk_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        t1.tofile(os.path.join(d, "r1.fq"))
        t2.tofile(os.path.join(d, "r2.fq"))
        env = dict(os.environ, RAYON_NUM_THREADS=str(threads))
        cmd = [exe, "--quiet", "-f", "extract", "-1", d + "/r1.fq", "-2", d + "/r2.fq", "-o", d + "/o1.fq", "-O",
               d + "/o2.fq", "-p", PATTERN1, "-e", str(MAX_ERROR)]
        try:
            total = 0.0
            for i in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                subprocess.run(cmd, check=True, env=env, capture_output=True)
                if i >= args.warmup:
                    total += time.perf_counter() - t0
            kind = "reference"
        except Exception as e:   # noqa: BLE001
            why = "reference binary failed (%s); C port instead" % e
        finally:
            shutil.rmtree(d, ignore_errors=True)
    if kind == "port":
        for _ in range(min(args.warmup, 2)):
            cpu_port_rate(t1, off, t2, off, n, threads)
        total = 0.0
        for _ in range(args.steps):
            _, dt, _ = cpu_port_rate(t1, off, t2, off, n, threads)
            total += dt
    value = args.steps * n / total
    sample = ("%d pairs per step (the CUDA arm's step), %s" %
              (n, "file -> file on tmpfs" if kind == "reference" else "pre-tokenised, host memory -> host memory"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n, "read_len": READ_LEN},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "reference_binary": why,
                         "note": "kind=port: oracle/cpu_extract.c, C restatement of the reference (Rust; no toolchain "
                                 "in this image) with ideal chunked parallelism = upper bound of its structure"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- helpers of the CUDA arm ---------------------------------------------------------------------------------------

def make_batch(_lib, ctx, cfg, first_idx, n, paired=True):
    mates = []
    for mate in ((1, 2) if paired else (1,)):
        d_text = ctx.dev_alloc(n * RL + 64)
        d_off = ctx.dev_alloc((n + 1) * 4)
        ctx.generate_device(cfg, mate, SEED, first_idx, n, d_text, d_off)
        m = _lib.MateIn()
        m.text, m.rec_off, m.text_len, m.max_rec_bytes = d_text, d_off, n * RL, RL
        mates.append(m)
    return mates


def free_batch(ctx, mates):
    for m in mates:
        ctx.dev_free(m.text)
        ctx.dev_free(m.rec_off)


def kernel_name(paired, one_pattern, general):
    if paired and one_pattern:
        # two kernels per step (bk_compact.cuh): the single-end extract kernel on the pattern's mate, then the ordered
        # compaction of the other mate; timed together
        return ("bk::extract_kernel<PAIRED=false, GENERAL=%s> on the pattern's mate + bk::compact_kernel on the other"
                % ("true" if general else "false"))
    if paired and general:
        # two patterns, one of them needs a search: match_kernel on that mate, then ONE paired extract launch in table
        # form (the searched mate's front warps read its matches from the table, the anchored mate matches for itself)
        # and a one-thread merge of the matchers' error bits; timed together
        return ("bk::match_kernel on the searched mate + bk::extract_kernel<PAIRED=true, GENERAL=false, TABLE=true> on both "
                "mates (+ merge_error_kernel)")
    return "bk::extract_kernel<PAIRED=%s, GENERAL=%s>" % ("true" if paired else "false", "true" if general else "false")


def other_configs(_lib, units, peak):
    """C1, C2, C4 (-e 0/1/2) and the device tokenizer, device-timed like the headline: 2 warm-up + 3 timed launches on
    `units` resident units each (fresh input per launch is >> L2)."""
    import numpy as np
    out = {}
    for name, ks in (("C1", (1,)), ("C2", (1,)), ("C4", (0, 1, 2)), ("F2_unanchored_3_variants", (1,)), ("F2_unbounded_greedy", (1,)),
                     ("F2_general_repeated_run", (1,))):
        p1, p2, paired, plant1, plant2 = OTHER_CONFIGS[name]
        cfg = _lib.syn_config(READ_LEN, plant1=plant1, plant2=plant2)
        for k in ks:
            ctx = _lib.Context(_lib.Program(p1, k) if p1 else None, _lib.Program(p2, k) if p2 else None, paired=paired)
            mates = make_batch(_lib, ctx, cfg, 0, units, paired)
            cap = units * (RL + 120)
            d_o1 = ctx.dev_alloc(cap)
            d_o2 = ctx.dev_alloc(cap) if paired else 0
            ms, launches = [], 0
            for i in range(5):
                r = ctx.extract_device(units, mates[0], mates[1] if paired else None, d_o1, cap, d_o2, cap if paired else 0)
                if i >= 2:
                    ms.append(r.kernel_ms)
                    launches += r.launches
            t = sum(ms) / len(ms)
            alg = len(mates) * units * RL + r.out1_len + r.out2_len
            general = name == "C4"
            out["%s_k%d" % (name, k)] = {
                "units": units, "unit": "pairs" if paired else "reads", "units_per_s": units / (t / 1e3),
                "ms_per_launch": t, "kept_fraction": r.n_out / units, "alg_bytes": alg,
                "achieved_gbs": alg / (t / 1e3) / 1e9, "frac": alg / (t / 1e3) / 1e9 / peak,
                "kernel": F2_KERNELS if name.startswith("F2") else kernel_name(paired, (p1 is None) != (p2 is None), general),
                "launches": launches}
            if name == "C1":   # the device tokenizer over the same text, wall time around the call (it ends with small reads)
                d_off = ctx.dev_alloc((units + 2) * 4)
                n, consumed, maxrec = C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
                best = None
                for it in range(4):
                    t0 = time.perf_counter()
                    rc = _lib.lib.bk_tokenize_device(ctx.handle, mates[0].text, units * RL, 1, d_off, units + 1, C.byref(n),
                                                     C.byref(consumed), C.byref(maxrec))
                    dt = time.perf_counter() - t0
                    if rc != 0 or n.value != units:
                        raise SystemExit("bk_tokenize_device failed in the bench: rc %d, %d records" % (rc, n.value))
                    best = dt if it and (best is None or dt < best) else (best if it else None)
                same = bool((ctx.d2h(d_off, (units + 1) * 4, np.uint32) == ctx.d2h(mates[0].rec_off, (units + 1) * 4, np.uint32)).all())
                out["tokenizer"] = {"units": units, "unit": "records", "units_per_s": units / best, "ms_per_call": best * 1e3,
                                    "alg_bytes": units * RL, "achieved_gbs": units * RL / best / 1e9,
                                    "frac": units * RL / best / 1e9 / peak, "kernel": "bk::tok_fused_kernel (+ tok_init_kernel)",
                                    "offsets_equal_generator": same,
                                    "note": "wall time around bk_tokenize_device incl. its result reads; text resident"}
                ctx.dev_free(d_off)
            free_batch(ctx, mates)
            ctx.dev_free(d_o1)
            if d_o2:
                ctx.dev_free(d_o2)
            ctx.close()
    return out


def crc_file(path):
    crc = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(64 << 20)
            if not b:
                return crc
            crc = zlib.crc32(b, crc)


def file_e2e(_lib, ctx, cfg, pairs, world):
    """The product's file -> file path (bk_run through the `barkit` CLI) on tmpfs, --gpus 1 and --gpus world; the CPU
    port on the same files beside it."""
    import numpy as np
    need = 4 * pairs * RL + (64 << 20)
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.path.isdir(cand) and shutil.disk_usage(cand).free > 1.3 * need:
                base = cand
                break
        except OSError:
            pass
    if base is None:
        return {"skipped": "no scratch file system with %.1f GB free" % (need / 1e9)}
    d = tempfile.mkdtemp(prefix="bk_bench_", dir=base)
    try:
        mates = make_batch(_lib, ctx, cfg, 1 << 40, pairs, True)
        for i, m in enumerate(mates):
            ctx.d2h(m.text, pairs * RL).tofile(os.path.join(d, "r%d.fq" % (i + 1)))
        free_batch(ctx, mates)
        res = {"pairs": pairs, "fs": base, "bytes_in": 2 * pairs * RL}
        crcs = {}
        for g in sorted({1, world}):
            best = None
            for rep in range(3):
                for f in ("o1.fq", "o2.fq"):
                    try:
                        os.remove(os.path.join(d, f))
                    except FileNotFoundError:
                        pass
                time.sleep(1.0)   # (the previous run's process is still being torn down by the driver)
                t0 = time.perf_counter()
                r = subprocess.run([CLI, "--quiet", "extract", "-1", d + "/r1.fq", "-2", d + "/r2.fq", "-o", d + "/o1.fq",
                                    "-O", d + "/o2.fq", "-p", PATTERN1, "-e", str(MAX_ERROR), "--gpus", str(g)],
                                   capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    return {"error": "barkit --gpus %d failed: %s" % (g, r.stderr.strip()[:300])}
                best = dt if best is None else min(best, dt)
            crcs[g] = (crc_file(d + "/o1.fq"), crc_file(d + "/o2.fq"))
            res["gpus_%d" % g] = {"wall_s": best, "pairs_per_s": pairs / best}
        res["bytes_out"] = os.path.getsize(d + "/o1.fq") + os.path.getsize(d + "/o2.fq")
        # what the file system itself takes: two new files written side by side, one thread each (write(2) holds the
        # file's lock, more threads per file add nothing: tools/fs_peak.py)
        try:
            from concurrent.futures import ThreadPoolExecutor
            blk = np.full(8 << 20, 65, np.uint8)
            per_file = min(res["bytes_out"] // 2, 2 << 30)

            def wr(i):
                fd = os.open(os.path.join(d, "fs%d" % i), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
                for o in range(0, per_file, len(blk)):
                    os.pwrite(fd, blk, o)
                os.close(fd)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(2) as ex:
                list(ex.map(wr, range(2)))
            dt = time.perf_counter() - t0
            res["fs_write_gbs_two_files"] = 2 * per_file / dt / 1e9
            res["fs_bound_pairs_per_s"] = pairs / (res["bytes_out"] / (2 * per_file / dt))
            for i in range(2):
                os.remove(os.path.join(d, "fs%d" % i))
        except Exception as e:   # noqa: BLE001
            res["fs_write_gbs_two_files"] = "failed: %s" % e
        res["gpus"] = world
        res["pairs_per_s"] = res["gpus_%d" % world]["pairs_per_s"]
        res["crc_equal_to_1gpu"] = (crcs[world] == crcs[1]) if world > 1 else None
        res["crc32"] = list(crcs[world])
        # the CPU port, file -> file: read both files, tokenise + extract on all cores, write both outputs
        try:
            from oracle import cport
            threads = host_threads()
            t0 = time.perf_counter()
            a1, a2 = np.fromfile(d + "/r1.fq", np.uint8), np.fromfile(d + "/r2.fq", np.uint8)
            r = cport.extract(a1, a2, PATTERN1, None, MAX_ERROR, threads=threads)
            with open(d + "/c1.fq", "wb") as f:
                f.write(r["out1"])
            with open(d + "/c2.fq", "wb") as f:
                f.write(r["out2"])
            dt = time.perf_counter() - t0
            res["cpu_port"] = {"wall_s": dt, "pairs_per_s": pairs / dt, "cores": threads,
                               "crc_equal": (zlib.crc32(r["out1"]), zlib.crc32(r["out2"])) == crcs[1]}
        except Exception as e:   # noqa: BLE001
            res["cpu_port"] = {"error": str(e)[:200]}
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


def measure_traffic(pairs):
    """DRAM bytes of one extract launch of THIS build on THIS box: a child run of this script under ncu with the two
    dram counters only.  -> (bytes per pair, source) or (None, why)."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    fd, path = tempfile.mkstemp(suffix=".csv")
    os.close(fd)
    try:
        # the second step's two launches (extract on R1, compaction of R2), serialised by ncu
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control",
               "none", "-k", "regex:extract_kernel|compact_kernel", "-s", "2", "-c", "2", "--csv", "--log-file", path,
               sys.executable, os.path.abspath(__file__), "--steps", "1", "--warmup", "1", "--pairs", str(pairs), "--child"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        tot = 0.0
        seen = 0
        shares = {}
        import csv
        for row in csv.reader(open(path)):
            if len(row) > 5 and row[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                v = float(row[-1].replace(",", ""))
                unit = row[-2].lower()
                kname = "compact_kernel" if "compact_kernel" in ",".join(row) else "extract_kernel"
                if row[-3] == "gpu__time_duration.sum":
                    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit.replace("second", "s").replace("nsecond", "ns"), 1e-6)
                    shares.setdefault(kname, {})["ms_under_ncu"] = v
                    continue
                v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1)
                shares.setdefault(kname, {})["dram_bytes"] = shares.get(kname, {}).get("dram_bytes", 0.0) + v
                tot += v
                seen += 1
        if seen == 4:
            return tot / pairs, "ncu capture of this run's build on this box (child process, %d pairs; both kernels of a step)" % pairs, shares
        return None, "ncu produced no counters (rc %d): %s" % (r.returncode, (r.stderr or r.stdout).strip()[-160:]), None
    except Exception as e:   # noqa: BLE001
        return None, "ncu failed: %s" % e, None
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per step (device-resident)")
    ap.add_argument("--e2e-pairs", type=int, default=2_000_000, help="read pairs per e2e step (host buffers)")
    ap.add_argument("--config-units", type=int, default=10_000_000, help="units per launch of the other configs")
    ap.add_argument("--file-pairs", type=int, default=10_000_000, help="read pairs of the file -> file leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-file-e2e", action="store_true")
    ap.add_argument("--no-traffic", action="store_true")
    ap.add_argument("--child", action="store_true", help="(internal) the kernel only: the run ncu captures for roofline.traffic")
    args = ap.parse_args()
    if args.child:
        args.no_cpu_baseline = args.no_e2e = args.no_configs = args.no_file_e2e = args.no_traffic = True

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

    def sum_over_ranks(x):
        if dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if _lib.device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device %d; the extract path has no CPU fallback" % local_rank)
    prog = _lib.Program(PATTERN1, MAX_ERROR)
    # one process per GPU: the host threads of the library (output of the pattern-less mate in e2e) are shared out
    # among the ranks of this box instead of every rank starting one per core
    lib_threads = max(2, host_threads() // max(world, 1))
    ctx = _lib.Context(prog, None, paired=True, device=local_rank, host_threads=lib_threads)
    cfg = _lib.syn_config(READ_LEN, plant1=(b"ATGCCAT", 16, 0))
    P = args.pairs
    nbatch = max(1, min(4, args.steps + args.warmup))

    # distinct batches per rank and per step slot: rank r owns records [r*2^36 + b*P, ...)
    batches = [make_batch(_lib, ctx, cfg, multigpu.rank_first_index(rank, b, P), P) for b in range(nbatch)]
    cap1 = ctx.out_bound(1, P, P * RL)
    cap2 = ctx.out_bound(2, P, P * RL)
    d_o1, d_o2 = ctx.dev_alloc(cap1), ctx.dev_alloc(cap2)

    def step(i):
        m = batches[i % nbatch]
        return ctx.extract_device(P, m[0], m[1], d_o1, cap1, d_o2, cap2)

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
        alg_bytes += 2 * P * RL + r.out1_len + r.out2_len
        kept += r.n_out
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    sampler.end()
    clocks = sampler.stop()
    dev_ms_max = max_over_ranks(dev_ms)
    value = world * args.steps * P / (dev_ms_max / 1e3)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (dev_ms / 1e3) / 1e9
    if args.child:
        ctx.close()
        return

    # ---- e2e through host buffers (pinned), double-buffered over the ctx's two slots ---------------
    e2e = None
    if not args.no_e2e:
        Q = min(args.e2e_pairs, P)
        nb = Q * RL
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
            h.text, h.rec_off, h.text_len, h.max_rec_bytes = hbuf[t][0], hbuf["off"][0], nb, RL
        outs = [("o1a", "o2a"), ("o1b", "o2b")]
        res = _lib.Result()

        def e2e_mode(c, nsteps, nwarm, raw=False):
            def submit(slot):
                if raw:   # raw FASTQ text in, record offsets made on the device (bk_submit_text)
                    rc = _lib.lib.bk_submit_text(c.handle, slot, Q, hbuf["t1"][0], nb, hbuf["t2"][0], nb, RL)
                else:
                    rc = _lib.lib.bk_submit(c.handle, slot, Q, C.byref(h1), C.byref(h2))
                if rc != 0:
                    raise SystemExit("bk_submit failed: %s" % _lib.lib.bk_last_error(c.handle).decode())

            def wait(slot):
                a, b = outs[slot]
                rc = _lib.lib.bk_wait(c.handle, slot, hbuf[a][0], hbuf[a][1], hbuf[b][0], hbuf[b][1], C.byref(res))
                if rc != 0:
                    raise SystemExit("bk_wait failed: %s" % _lib.lib.bk_last_error(c.handle).decode())
                return (res.h2d_bytes, res.d2h_bytes, res.launches)

            def run(k):
                last, nl = (0, 0, 0), 0
                submit(0)
                for s in range(k):
                    if s + 1 < k:
                        submit((s + 1) & 1)
                    last = wait(s & 1)
                    nl += last[2]
                return last, nl

            run(nwarm)
            barrier()
            t0 = time.perf_counter()
            last, nl = run(nsteps)
            barrier()
            secs = max_over_ranks(time.perf_counter() - t0)
            return world * nsteps * Q / secs, secs, last, nl

        v_auto, s_auto, last, e2e_launches = e2e_mode(ctx, args.steps, max(2, min(args.warmup, 3)))
        modes = {}
        short = max(3, args.steps)   # (same length as the default mode: a short run is dominated by filling and draining the two slots)
        for name, flag in (("host_mate_pinned", _lib.FLAG_HOST_MATE_PINNED), ("device_both_mates", _lib.FLAG_DEVICE_BOTH_MATES)):
            c2 = _lib.Context(prog, None, paired=True, device=local_rank, extra_flags=flag, host_threads=lib_threads)
            v, _, l2, _ = e2e_mode(c2, short, 2)
            modes[name] = {"value": v, "h2d_bytes_per_step": int(l2[0]), "d2h_bytes_per_step": int(l2[1])}
            if name == "device_both_mates":
                v, _, l2, _ = e2e_mode(c2, short, 2, raw=True)
                modes["raw_text_device_tokenizer"] = {"value": v, "h2d_bytes_per_step": int(l2[0]), "d2h_bytes_per_step": int(l2[1])}
            c2.close()
        # the box's host-link ceiling with every rank's GPU copying both ways at once
        barrier()
        up, down, duplex = ctx.bus_probe(256 << 20, 6)
        barrier()
        bus_sum = sum_over_ranks(duplex)
        moved = (last[0] + last[1]) * world * args.steps / s_auto / 1e9     # GB/s that crossed the links in the default mode
        for mname, md in modes.items():   # ... and in every pinned mode (all ranks): pairs/s x bytes per pair over the links
            md["bus_used_gbs"] = md["value"] * (md["h2d_bytes_per_step"] + md["d2h_bytes_per_step"]) / Q / 1e9
            md["bus_frac"] = md["bus_used_gbs"] / bus_sum if bus_sum else None
        # The host's memory system: every byte of a pair's input is read from host memory and every byte of its output
        # written to it exactly once, whichever side does it (DMA engine, or the cores assembling the host-side mate).
        # What all host cores move with plain copies, measured here with the GPUs idle, bounds that.
        host_mem_peak = None
        if rank == 0:
            g = C.c_double(0.0)
            if _lib.lib.bk_host_mem_probe(128 << 20, host_threads(), 3, C.byref(g)) == 0:
                host_mem_peak = g.value
        barrier()
        bytes_in_out = 2 * nb + int(res.out1_len) + int(res.out2_len)      # per step: both mates in, both outputs out
        host_mem_used = v_auto * bytes_in_out / Q / 1e9
        e2e = {"value": v_auto, "unit": UNIT,
               "h2d_bytes_per_step": int(last[0]), "d2h_bytes_per_step": int(last[1]),
               "pairs_per_step": Q, "ms_per_step": 1e3 * s_auto / args.steps, "host_threads": lib_threads,
               "modes": modes,
               "bus_peak_gbs": bus_sum, "bus_h2d_gbs_rank0": up, "bus_d2h_gbs_rank0": down,
               "bus_used_gbs": moved, "bus_frac": moved / bus_sum if bus_sum else None,
               "host_mem_peak_gbs": host_mem_peak, "host_mem_used_gbs": host_mem_used,
               "host_mem_frac": host_mem_used / host_mem_peak if host_mem_peak else None,
               "limiter": ("host memory system" if host_mem_peak and host_mem_used / host_mem_peak >= (moved / bus_sum if bus_sum else 0)
                           else "host links"),
               "note": "bk_submit/bk_wait with pinned host buffers, 2 slots in flight; h2d/d2h bytes as the "
                       "library counts them (text + record offsets in, FASTQ text + result + match table out). "
                       "value = the library's default: R2 has no pattern in this config, so it stays on the host and "
                       "its output is assembled there from the input buffer inside the timed region, unless the context "
                       "measures that to be slower than the bus. modes = the two choices pinned, and raw_text_device_tokenizer = "
                       "bk_submit_text: no offsets from the host, both mates tokenised on the device in front of the extract "
                       "kernels (what the `barkit` command line uses). bus_peak_gbs = pinned "
                       "H2D + D2H at once on every rank's GPU (sum over ranks), measured in this run. host_mem_peak_gbs = "
                       "read + written bytes per second of plain copies on all host cores with the GPUs idle (bk_host_mem_probe); "
                       "host_mem_used_gbs = the default mode's pairs/s x (input + output bytes of a pair), each of which crosses "
                       "host memory once whichever side moves it; limiter = the larger of host_mem_frac and bus_frac."}
    else:
        e2e_launches = 0

    # ---- CPU baseline beside it (rank 0; bounded sample) -------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = host_threads()
        probe = min(200_000, P)
        budget = 12.0 if world == 1 else 5.0
        m = batches[0]
        t1 = ctx.d2h(m[0].text, min(P, 6_000_000) * RL)
        t2 = ctx.d2h(m[1].text, min(P, 6_000_000) * RL)
        off = (np.arange(min(P, 6_000_000) + 1, dtype=np.uint64) * RL).astype(np.uint32)
        rate, _, _ = cpu_port_rate(t1, off, t2, off, probe, threads)
        n = int(min(max(rate * budget, probe), min(P, 6_000_000)))
        _, dt1, _ = cpu_port_rate(t1, off, t2, off, n, threads)
        reps = int(max(1, min(200, round(budget / max(dt1, 1e-3)))))
        total = 0.0
        for _ in range(reps):
            _, dt, r = cpu_port_rate(t1, off, t2, off, n, threads)
            total += dt
        rate = reps * n / total
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d passes over the first %d pairs of step 0's batch (%.1f s), pre-tokenised, "
                         "host memory -> host memory" % (reps, n, total),
               "reference_binary": reference_binary()[1],
               "note": "oracle/cpu_extract.c, C restatement of the reference (Rust toolchain absent); "
                       "ideal chunked parallelism, so an upper bound of the reference's structure"}
        del t1, t2
    if world > 1:
        barrier()

    # ---- every other BASELINE config, device-timed (one GPU) ----------------------------------------------
    for b in batches[1:]:
        free_batch(ctx, b)
    batches = batches[:1]
    configs = None
    if world == 1 and not args.no_configs:
        ctx.dev_free(d_o1)
        ctx.dev_free(d_o2)
        d_o1 = d_o2 = 0
        configs = other_configs(_lib, min(args.config_units, 12_000_000), peak)

    # ---- the product's file -> file path at --gpus world (rank 0 drives the CLI; the others get out of the way) ------
    fe = None
    flag = os.path.join(tempfile.gettempdir(), "bk_bench_done_%s" % os.environ.get("MASTER_PORT", "0"))
    if not args.no_file_e2e:
        if rank == 0:
            try:
                os.remove(flag)
            except OSError:
                pass
        if world > 1:
            barrier()
        if rank == 0:
            fe = file_e2e(_lib, ctx, cfg, args.file_pairs, world)
            if world > 1:
                open(flag, "w").close()
        else:   # no collective here: a pending NCCL kernel would sit on this GPU while the CLI uses it
            t0 = time.time()
            while not os.path.exists(flag) and time.time() - t0 < 600:
                time.sleep(0.05)

    if rank == 0:
        traffic, traffic_src, shares = None, None, None
        if world == 1 and not args.no_traffic:
            tp = min(P, 2_000_000)
            per_pair, traffic_src, shares = measure_traffic(tp)
            if per_pair is not None:
                traffic = per_pair * P
        if traffic is None:
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    tj = json.load(f)
                traffic = tj["dram_bytes_per_pair"] * P
                traffic_src = "static: profiles/traffic.json (%s)%s" % (tj.get("source", "committed ncu capture"),
                                                                        "; live capture: " + traffic_src if traffic_src else "")
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": P, "read_len": READ_LEN,
                       "kept_fraction": kept / float(args.steps * P),
                       "l2": "each step reads %.1f GB of fresh input (>> 126 MB L2); %d distinct resident "
                             "batches cycled; no explicit flush" % (2 * P * RL / 1e9, nbatch),
                       "wall_ms_per_step": wall_ms / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_step": alg_bytes / args.steps,
                         "kernel": kernel_name(True, True, False) + " (2 launches per step, timed together)",
                         "kernel_shares": shares},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "configs": configs,
            "file_e2e": fe,
            "gpu_launches": launches,
            "gpu_launches_e2e": e2e_launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
