#!/bin/bash
set -u
OUT=gpurun_out/r02ab4
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -x -q -k "decode or pdl or golden or edge" > "$OUT/pytest.log" 2>&1; tail -3 "$OUT/pytest.log"
for rep in 1 2; do
  for lib in libflute_b200_ab_f4b86b2.so libflute_b200.so; do
    echo "== $lib rep $rep" | tee -a "$OUT/ab.log"
    FLUTE_B200_PY_OPS=1 FLUTE_B200_LIB=$lib timeout 200 python tools/microbench.py --M 1 --shapes llama8b >> "$OUT/ab.log" 2>&1
  done
done
grep -E "==|W4" "$OUT/ab.log"
timeout 300 python bench.py --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
