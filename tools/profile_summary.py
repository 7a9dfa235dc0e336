"""Key counters of every kernel in a set of ncu reports, as markdown (what profiles/<tag>_summary.md is made from).
  python tools/profile_summary.py a.ncu-rep [b.ncu-rep ...]      # -> stdout; raw / details pages are written beside it
  with --save <prefix>: also writes profiles/<prefix>_<report>_raw.csv and _details.txt"""
import csv, os, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
args = sys.argv[1:]
save = None
if "--save" in args:
    i = args.index("--save"); save = args[i + 1]; del args[i:i + 2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in args:
    name = os.path.splitext(os.path.basename(rep))[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    if save:
        short = name.split("_", 1)[1] if "_" in name else name
        open(os.path.join(root, "profiles", "%s_%s_raw.csv" % (save, short)), "w").write(raw)
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(root, "profiles", "%s_%s_details.txt" % (save, short)), "w").write(det)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("### %s: `%s`\n" % (name, r[kn]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("- %s = %s %s" % (k, r[i], units[i]))
        print()
