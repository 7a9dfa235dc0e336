"""Turn the round's ncu artefacts (gpurun_out/) into the tracked summaries under profiles/.
  python tools/make_profile_summary.py <tag> <launches.csv> <full.ncu-rep> "<bench command>" <pairs per launch>"""
import csv, json, os, subprocess, sys, collections
tag, launches, rep, cmd, pairs = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
# ---- launch list -------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if r[ix["Metric Name"]] == "gpu__time_duration.sum":
        v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        per[r[ix["Kernel Name"]]].append(v)
tot = sum(sum(v) for v in per.values())
with open(os.path.join(prof, tag + "_launches.csv"), "w") as f: f.write(open(launches).read())
# ---- full capture ------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(prof, tag + "_extract_kernel_full_raw.csv"), "w").write(raw)
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(os.path.join(prof, tag + "_extract_kernel_full_details.txt"), "w").write(det)
rr = list(csv.reader(raw.splitlines()))
m = {h: (v, u) for h, u, v in zip(rr[0], rr[1], rr[2])}
def val(k):
    v, u = m[k]; v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg"]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
json.dump({"source": "profiles/%s_extract_kernel_full_raw.csv (ncu --set full, one extract_kernel<true> launch, %d pairs)" % (tag, pairs),
           "dram_bytes_read": rd, "dram_bytes_write": wr, "pairs": pairs, "dram_bytes_per_pair": (rd + wr) / pairs},
          open(os.path.join(prof, "traffic.json"), "w"), indent=1)
with open(os.path.join(prof, tag + "_summary.md"), "w") as f:
    f.write("# %s ncu summary (B200, `%s`)\n\n" % (tag, cmd))
    f.write("Per-launch device times (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: compare shares):\n\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write("- `%s`: %d launches, mean %.3f ms, share of captured GPU time %.1f%%\n" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    f.write("\nThe bksyn-v1 generator launches are bench set-up (outside the timed region); inside the timed region the only kernel is "
            "`extract_kernel<true, false>` (paired, anchored patterns) (one launch per step, plus one cudaMemsetAsync of the look-back scratch).\n\n")
    f.write("Full capture of one `extract_kernel<true, false>` (paired, anchored patterns) launch (%d pairs, `ncu --set full --clock-control none`; file %s_extract_kernel_full_raw.csv):\n\n" % (pairs, tag))
    for k in keys:
        if k in m: f.write("- %s = %s %s\n" % (k, m[k][0], m[k][1]))
    f.write("\nDRAM traffic per launch = %.3f GB read + %.3f GB written = %.3f GB (%.1f B/pair).\n" % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, (rd + wr) / pairs))
print(open(os.path.join(prof, tag + "_summary.md")).read())
