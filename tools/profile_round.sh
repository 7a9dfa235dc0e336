#!/bin/bash
# Round profile: default bench (no profiler), ncu launch list of the same command, one full capture of the extract kernel.
set -x
python bench.py > gpurun_out/r1e_bench.log 2>&1; tail -1 gpurun_out/r1e_bench.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1e_bench_ref.log 2>&1; tail -1 gpurun_out/r1e_bench_ref.log | cut -c1-300
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1e_ncu1.log 2>&1
# (BK_KERNEL=4: the kernel the context's own timing picks for this workload, without the two trial launches of extract_kernel5)
BK_KERNEL=4 ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 3 -c 1 -o gpurun_out/r1e_extract_full -f \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1e_ncu2.log 2>&1
ls -la gpurun_out | tail -8
