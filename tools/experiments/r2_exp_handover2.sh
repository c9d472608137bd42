#!/bin/bash
# Hand-over A/B at M = 2, 4, 16 (previous build pinned to the decode kernel with variant 2) + per-CTA trace of the new build.
set -u
OUT=gpurun_out/r02h3
mkdir -p "$OUT"
for lib in libflute_b200_ab_base.so libflute_b200.so; do
  echo "== $lib"
  FLUTE_B200_PY_OPS=1 FLUTE_B200_LIB=$lib timeout 200 python tools/microbench.py --M 2,4,16 --variant 2 --shapes llama8b 2>&1 | grep "N="
done > "$OUT/ab_m.log" 2>&1
cut -c1-110 "$OUT/ab_m.log"
FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 1 --shapes llama8b --trace 2 > "$OUT/trace_m1.log" 2>&1
FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 4 --shapes llama8b --trace 2 > "$OUT/trace_m4.log" 2>&1
head -60 "$OUT/trace_m1.log" | cut -c1-200
