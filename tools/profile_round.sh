#!/bin/bash
# Round profile (one gpurun call): default bench and reference arm without a profiler, then -- each only after the same
# command has exited 0 without ncu -- the ncu launch list of the bench's timed path and one full capture of a step's two
# kernels (single-end extract of R1 + compaction of R2).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2}
set -x
python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err; tail -1 gpurun_out/${tag}_bench.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2>&1; tail -1 gpurun_out/${tag}_bench_ref.log | cut -c1-300
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-file-e2e --no-traffic"
$B > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"extract_kernel|compact_kernel" -s 6 -c 2 -o gpurun_out/${tag}_step_full -f $B > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out | tail -8
