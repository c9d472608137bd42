#!/bin/bash
# Round-2 experiment 9: split-K hand-over through owner-polled {value, 1} word slots (no reductions / counters).
set -u
OUT=gpurun_out/r02j
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -6 "$OUT/pytest_gpu.log" | cut -c1-250
timeout 200 python tools/microbench.py --M 1 --shapes llama8b > "$OUT/microbench_M1.log" 2>&1
cat "$OUT/microbench_M1.log"
timeout 300 python bench.py --no-configs --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"
python - <<'PY'
import json
for l in open('gpurun_out/r02j/bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
PY
for sh in small gateup; do
  echo "== trace $sh" >> "$OUT/trace.log"
  FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 1 --shapes $sh --trace 2 --reps 3 >> "$OUT/trace.log" 2>&1
done
grep -v "producer\|mma wait\|mma issue\|dq5" "$OUT/trace.log" | tail -50
