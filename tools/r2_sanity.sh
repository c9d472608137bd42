#!/bin/bash
# Shortest useful GPU check: the GPU suite and the smoke entry point.
set -u
OUT=gpurun_out/r02s
mkdir -p "$OUT"
timeout 120 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -8 "$OUT/pytest_gpu.log" | cut -c1-300
timeout 40 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; tail -4 "$OUT/smoke.log"
