#!/bin/bash
# ncu captures of the kernels next to the bench's: the device tokenizer, and C4's two launches (match_kernel on the searched
# mate + the paired extract launch in table form); plus C4's launch list; match_kernel<VM> on a general-form pattern.
tag=${1:-r2}
set -x
python tools/tokenize_speed.py > gpurun_out/${tag}_tok_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tok_fused -s 2 -c 1 -o gpurun_out/${tag}_tok_full -f python tools/tokenize_speed.py > gpurun_out/${tag}_tok_ncu.log 2>&1
BK_ONLY=C4 python tools/config_speed.py 10000000 > gpurun_out/${tag}_c4_plain.log 2>&1 && \
BK_ONLY=C4 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_c4_launches.csv python tools/config_speed.py 10000000 > gpurun_out/${tag}_c4_ncu1.log 2>&1
BK_ONLY=C4 ncu --set full --clock-control none --import-source on -k regex:"match_kernel|extract_kernel" -s 6 -c 2 -o gpurun_out/${tag}_c4_full -f python tools/config_speed.py 10000000 > gpurun_out/${tag}_c4_ncu.log 2>&1
BK_ONLY="general anchored" python tools/match_speed.py 4000000 > gpurun_out/${tag}_vm_plain.log 2>&1 && \
BK_ONLY="general anchored" ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 2 -c 1 -o gpurun_out/${tag}_vm_full -f python tools/match_speed.py 4000000 > gpurun_out/${tag}_vm_ncu.log 2>&1
ls -la gpurun_out | tail -6
