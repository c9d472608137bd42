#!/bin/bash
# Round-2 experiment 7: binding error path, full GPU suite, decode defaults (plain loop + 4 spare SMs), prefill with two dequant sets.
set -u
OUT=gpurun_out/r02g
mkdir -p "$OUT"
timeout 300 python -m pytest tests -m gpu -x -q -k "python_api_shapes or opcheck or binding" > "$OUT/pytest_api.log" 2>&1
echo "pytest api exit $?" >> "$OUT/pytest_api.log"; tail -4 "$OUT/pytest_api.log"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -6 "$OUT/pytest_gpu.log"
timeout 200 python tools/microbench.py --M 1 --shapes llama8b > "$OUT/microbench_M1.log" 2>&1
cat "$OUT/microbench_M1.log"
timeout 300 python tools/microbench.py --M 512,4096 --shapes llama8b > "$OUT/microbench_prefill.log" 2>&1
cat "$OUT/microbench_prefill.log"
timeout 300 python bench.py --no-configs --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"
python - <<'PY'
import json
for l in open('gpurun_out/r02g/bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
PY
FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 4096 --shapes gateup --trace 1 --reps 1 > "$OUT/trace_prefill.log" 2>&1
grep -v "^     start\|^     setup\|exit" "$OUT/trace_prefill.log" | tail -24
