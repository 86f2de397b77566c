#!/usr/bin/env bash
# Round-2 check S: filter view + relatedness-prune chaining through the device commands (tools/check_r2s.py).
mkdir -p gpurun_out
timeout 170 python tools/check_r2s.py 2>&1 | tail -40
