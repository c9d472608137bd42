#!/bin/bash
# Round-2 experiment 3: full footprint with separate activation barrier + entry L2 prefetch; per-CTA tail analysis.
set -u
OUT=gpurun_out/r02d
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or footprint or pdl or binding" > "$OUT/pytest_decode.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_decode.log"; tail -5 "$OUT/pytest_decode.log"
echo "== variant 4" | tee -a "$OUT/microbench_M1.log"
timeout 200 python tools/microbench.py --M 1 --shapes llama8b --variant -1 >> "$OUT/microbench_M1.log" 2>&1
cat "$OUT/microbench_M1.log"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline >> "$OUT/bench.log" 2>> "$OUT/bench.err"
cat "$OUT/bench.log"; tail -3 "$OUT/bench.err"
for sh in small gateup; do
  echo "== trace variant 4 $sh" >> "$OUT/trace.log"
  FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 1 --shapes $sh --variant -1 --trace 2 --reps 3 >> "$OUT/trace.log" 2>&1
done
cat "$OUT/trace.log"
