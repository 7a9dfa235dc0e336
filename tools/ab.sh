#!/bin/bash
# A/B of build-time variants on the box: tools/ab.sh "<defines A>" "<defines B>" ... ; each is built and timed with
# $BK_AB_CMD (default: tools/config_speed.py with BK_ONLY / BK_PAIRS from the environment).
cmd=${BK_AB_CMD:-python tools/config_speed.py ${BK_PAIRS:-10000000}}
for d in "$@"; do
  BK_DEFINES="$d" python -m barkit_b200.build --force >/dev/null 2>&1 || { echo "build failed: $d"; continue; }
  echo "== [$d]"; timeout -k 5 200 $cmd 2>&1 | tail -${BK_LINES:-6}
done
python -m barkit_b200.build --force >/dev/null 2>&1
