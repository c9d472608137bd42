#!/bin/bash
# Round-2 experiment 6: full GPU suite, dq-loop A/B, prefill after the hybrid schedule, full bench with configs.
set -u
OUT=gpurun_out/r02f
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -6 "$OUT/pytest_gpu.log"
for cfg in "0 0" "8 0" "0 4" "8 4"; do
  set -- $cfg
  echo "== ablate $1 slack $2" | tee -a "$OUT/microbench_M1.log"
  timeout 200 python tools/microbench.py --M 1 --shapes llama8b --ablate $1 --slack $2 >> "$OUT/microbench_M1.log" 2>&1
done
cat "$OUT/microbench_M1.log"
timeout 300 python tools/microbench.py --M 16,512,4096 --shapes llama8b > "$OUT/microbench_prefill.log" 2>&1
cat "$OUT/microbench_prefill.log"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -5 "$OUT/bench.err"
python - <<'PY'
import json
for l in open('gpurun_out/r02f/bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d.get('configs_error'))
        for c in d.get('configs', []):
            if 'shapes' in c:
                print(' ', c['name'], c['bound'], round(c['frac_min'],3), round(c['frac_max'],3), c['kernel'])
            else:
                print(' ', c['name'], round(c['value'],1), 'tok/s', round(c['roofline']['frac'],3), c['kernel'])
PY
