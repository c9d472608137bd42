#!/bin/bash
# Round-2 experiment 2: co-resident footprint with 16 dequant warps / 40 registers / blocking TMEM hand-over.
set -u
OUT=gpurun_out/r02b
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or footprint or pdl or python_api or opcheck" > "$OUT/pytest_decode.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_decode.log"; tail -5 "$OUT/pytest_decode.log"
for cfg in "4 -1" "3 0" "3 6" "3 12"; do
  set -- $cfg
  echo "== variant $1 l2pf $2" | tee -a "$OUT/microbench_M1.log"
  timeout 200 python tools/microbench.py --M 1 --shapes llama8b --variant $1 --l2pf $2 >> "$OUT/microbench_M1.log" 2>&1
done
cat "$OUT/microbench_M1.log"
for v in 4 3 $((3 | (1 << 16))) $((3 | (13 << 16))); do
  echo "== bench variant $v" | tee -a "$OUT/bench.log"
  FLUTE_B200_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline >> "$OUT/bench.log" 2>> "$OUT/bench.err"
done
cat "$OUT/bench.log"; tail -3 "$OUT/bench.err"
for v in 4 3; do
  for sh in small gateup; do
    echo "== trace variant $v $sh" >> "$OUT/trace.log"
    FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 1 --shapes $sh --variant $v --trace 1 --reps 3 >> "$OUT/trace.log" 2>&1
  done
done
cat "$OUT/trace.log"
