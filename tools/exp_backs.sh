#!/bin/bash
# A/B of the single-end kernel's back-warp count on the box: speed on C1 / C3, then a parity subset per variant
K='baseline_configs_small or ragged or runs_of_unchanged or derived or crlf or variable_repetition_single or unbounded_repetition_single or empty_and_tiny or long_records or device_resident'
for b in "$@"; do
  BK_DEFINES="BK_SE_BACKS=$b" python -m barkit_b200.build --force >/dev/null 2>&1 || { echo "build failed: $b"; continue; }
  echo "== BK_SE_BACKS=$b"
  BK_ONLY=C1 timeout -k 5 120 python tools/config_speed.py 10000000 2>&1 | tail -1
  BK_ONLY=C3 timeout -k 5 120 python tools/config_speed.py 10000000 2>&1 | tail -1
  BK_ONLY=C2 timeout -k 5 120 python tools/config_speed.py 10000000 2>&1 | tail -1
  [ "$b" != 4 ] && timeout -k 5 300 python -m pytest tests/test_extract_gpu.py -m gpu -x -q -k "$K" 2>&1 | tail -3
done
python -m barkit_b200.build --force >/dev/null 2>&1
