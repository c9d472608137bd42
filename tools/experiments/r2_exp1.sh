#!/bin/bash
# Round-2 experiment 1: half-SM decode footprint (co-resident consecutive kernels under PDL) vs full-SM.
set -u
OUT=gpurun_out/r02a
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or footprint or pdl" > "$OUT/pytest_decode.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_decode.log"; tail -5 "$OUT/pytest_decode.log"
for cfg in "-1 -1" "4 -1" "3 0" "3 3" "3 6" "3 12"; do
  set -- $cfg
  echo "== variant $1 l2pf $2" | tee -a "$OUT/microbench_M1.log"
  timeout 200 python tools/microbench.py --M 1 --shapes llama8b --variant $1 --l2pf $2 >> "$OUT/microbench_M1.log" 2>&1
done
echo "== variant 3 l2pf 6 grid 296" | tee -a "$OUT/microbench_M1.log"
timeout 200 python tools/microbench.py --M 1 --shapes llama8b --variant 3 --l2pf 6 --force-grid 296 >> "$OUT/microbench_M1.log" 2>&1
cat "$OUT/microbench_M1.log"
for v in -1 3; do
  echo "== bench variant $v" | tee -a "$OUT/bench.log"
  FLUTE_B200_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline >> "$OUT/bench.log" 2>> "$OUT/bench.err"
done
FLUTE_B200_VARIANT=$((3 | (13 << 16))) timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline >> "$OUT/bench.log" 2>> "$OUT/bench.err"
FLUTE_B200_VARIANT=$((3 | (1 << 16))) timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline >> "$OUT/bench.log" 2>> "$OUT/bench.err"
cat "$OUT/bench.log"; tail -3 "$OUT/bench.err"
# timelines (profiling build): 3 chained launches
for v in 4 3; do
  for sh in small gateup; do
    echo "== trace variant $v $sh" >> "$OUT/trace.log"
    FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 1 --shapes $sh --variant $v --trace 1 --reps 3 >> "$OUT/trace.log" 2>&1
  done
done
cat "$OUT/trace.log"
