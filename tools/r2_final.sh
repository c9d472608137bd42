#!/bin/bash
# Round-2 closing run on one B200: the driver's own sequence (GPU suite, smoke, bench) + per-shape microbench.
set -u
OUT=gpurun_out/r02z
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -6 "$OUT/pytest_gpu.log" | cut -c1-250
timeout 300 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; tail -5 "$OUT/smoke.log"
timeout 300 python tools/microbench.py --M 1,16,512,4096 --shapes llama8b --json "$OUT/microbench_llama8b.json" > "$OUT/microbench_llama8b.log" 2>&1
cat "$OUT/microbench_llama8b.log"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -3 "$OUT/bench.err"
python - <<'PY'
import json
for l in open('gpurun_out/r02z/bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d.get('configs_error'), d['clocks'])
        for c in d.get('configs', []):
            if 'shapes' in c:
                print(' ', c['name'], c['bound'], round(c['frac_min'],3), round(c['frac_max'],3), c['kernel'])
            else:
                print(' ', c['name'], round(c['value'],1), 'tok/s', round(c['roofline']['frac'],3), c['kernel'])
        print(d.get('cpu_baseline'))
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; cat "$OUT/bench_reference.json" | cut -c1-400
# 3-bit decode: per-shape time and role cycles of the general kernel (profiling build)
FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --bits 3 --dtype fp16 --M 1 --shapes llama8b --trace 1 > "$OUT/w3_m1_profile.log" 2>&1
cut -c1-150 "$OUT/w3_m1_profile.log" | head -120
