#!/bin/bash
# Quick GPU loop: parity tests, a short bench, and the instruction-cache counters of one launch.
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 150 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value']/1e9, 'Gpairs/s', d['roofline']['frac'])"
timeout 200 ncu --clock-control none -k regex:extract_kernel -c 1 --metrics sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed python bench.py --steps 1 --warmup 1 --pairs 4000000 --no-e2e --no-cpu-baseline 2>&1 | grep -E "icc|gcc|inst_executed|time_duration|wavefronts"
