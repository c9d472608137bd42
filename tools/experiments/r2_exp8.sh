#!/bin/bash
set -u
OUT=gpurun_out/r02i
mkdir -p "$OUT"
python tools/binding_errpath.py > "$OUT/binding_errpath.log" 2>&1; cat "$OUT/binding_errpath.log"
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -12 "$OUT/pytest_gpu.log" | cut -c1-300
