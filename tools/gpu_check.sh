#!/bin/bash
# Quick GPU loop: smoke, device-timed speed of every BASELINE config, parity tests.  Every step under its own
# timeout so that a hung kernel cannot hold the box.
mkdir -p gpurun_out
echo "== smoke"; timeout -k 5 180 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== config speed (4 M units, device-timed)"; timeout -k 5 300 python tools/config_speed.py 4000000 2>&1 | tail -8
echo "== gpu tests"; timeout -k 5 ${1:-900} python -m pytest tests -m gpu -x -q 2>&1 | tail -15
