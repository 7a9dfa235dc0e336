#!/bin/bash
# A/B of build-time variants on the box: per variant the device-timed speed of C1 / C2 / C3 / C4 (k = 1, 10 M units) and
# a parity subset of the GPU tests.   tools/exp_defs.sh "<defines A>" "<defines B>" ...   (BK_NOTEST=1: speeds only)
K='baseline_configs_small or ragged or runs_of_unchanged or derived or crlf or variable_repetition_single or unbounded_repetition_single or empty_and_tiny or long_records or device_resident or pe_policy or readme'
for d in "$@"; do
  BK_DEFINES="$d" python -m barkit_b200.build --force >/dev/null 2>&1 || { echo "build failed: $d"; continue; }
  echo "== [$d]"
  for c in ${BK_CFGS:-C1 C3 C2 C4}; do BK_ONLY=$c timeout -k 5 120 python tools/config_speed.py 10000000 2>&1 | tail -1; done
  [ -z "$BK_NOTEST" ] && timeout -k 5 300 python -m pytest tests/test_extract_gpu.py -m gpu -x -q -k "$K" 2>&1 | tail -3
done
python -m barkit_b200.build --force >/dev/null 2>&1
