#!/bin/bash
# A/B of build-time variants on the box: tools/ab.sh "<defines A>" "<defines B>" ... ; each is built and timed with
# tools/config_speed.py (BK_ONLY / BK_PAIRS from the environment).
for d in "$@"; do
  BK_DEFINES="$d" python -m barkit_b200.build --force >/dev/null 2>&1 || { echo "build failed: $d"; continue; }
  echo "== [$d]"; timeout -k 5 200 python tools/config_speed.py ${BK_PAIRS:-10000000} 2>&1 | tail -${BK_LINES:-6}
done
python -m barkit_b200.build --force >/dev/null 2>&1
