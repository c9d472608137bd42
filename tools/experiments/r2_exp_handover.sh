#!/bin/bash
# Packed-atomic split-K hand-over: GPU suite, then same-box A/B against the previous build, then the bench.
set -u
OUT=gpurun_out/r02h2
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -6 "$OUT/pytest_gpu.log" | cut -c1-250
for rep in 1 2; do
  for lib in libflute_b200_ab_base.so libflute_b200.so; do
    echo "== $lib rep $rep"
    FLUTE_B200_PY_OPS=1 FLUTE_B200_LIB=$lib timeout 200 python tools/microbench.py --M 1 --shapes llama8b 2>&1 | grep "N="
  done
done > "$OUT/ab.log" 2>&1
cat "$OUT/ab.log" | cut -c1-110
timeout 200 python tools/microbench.py --M 2,4,16 --shapes llama8b 2>&1 | grep "N=" | cut -c1-110 | tee "$OUT/m_small.log"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -3 "$OUT/bench.err"
python - <<'PY'
import json
for l in open('gpurun_out/r02h2/bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d.get('configs_error'), d['clocks'])
        for c in d.get('configs', []):
            if 'shapes' in c:
                print(' ', c['name'], c['bound'], round(c['frac_min'],3), round(c['frac_max'],3), c['kernel'])
            else:
                print(' ', c['name'], round(c['value'],1), 'tok/s', round(c['roofline']['frac'],3), c['kernel'])
PY
