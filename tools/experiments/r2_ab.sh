#!/bin/bash
# Same-box A/B of library builds (FLUTE_B200_LIB): per-shape decode microbench, two repetitions each.
set -u
OUT=gpurun_out/r02ab5
mkdir -p "$OUT"
for rep in 1 2; do
  for lib in libflute_b200.so libflute_b200_ab_v1.so libflute_b200_ab_v2.so libflute_b200_ab_v3.so; do
    echo "== $lib rep $rep" | tee -a "$OUT/ab.log"
    FLUTE_B200_PY_OPS=1 FLUTE_B200_LIB=$lib timeout 200 python tools/microbench.py --M 1 --shapes llama8b >> "$OUT/ab.log" 2>&1
  done
done
grep -E "==|W4" "$OUT/ab.log"
