#!/bin/bash
# Round-2 experiment 5: pipelined dequant loop, SM slack sweep, stages=2 failure repro.
set -u
OUT=gpurun_out/r02e
mkdir -p "$OUT"
timeout 600 python tools/repro_sched.py > "$OUT/repro_sched.log" 2>&1
cat "$OUT/repro_sched.log"
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or pdl or binding" --deselect tests/test_qgemm_gpu.py::test_decode_kernel_schedules > "$OUT/pytest_decode.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_decode.log"; tail -4 "$OUT/pytest_decode.log"
for slack in 0 4 8 12 20; do
  echo "== slack $slack" | tee -a "$OUT/microbench_M1.log"
  timeout 200 python tools/microbench.py --M 1 --shapes llama8b --slack $slack >> "$OUT/microbench_M1.log" 2>&1
done
cat "$OUT/microbench_M1.log"
for slack in 0 4 8 12 20; do
  echo "== bench slack $slack" >> "$OUT/bench.log"
  FLUTE_B200_VARIANT=$((255 | (slack << 24))) timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-configs >> "$OUT/bench.log" 2>> "$OUT/bench.err"
done
python - <<'PY'
import json
for l in open('gpurun_out/r02e/bench.log'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
    else: print(l)
PY
tail -3 "$OUT/bench.err"
for sh in small gateup; do
  echo "== trace $sh" >> "$OUT/trace.log"
  FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 1 --shapes $sh --trace 2 --reps 3 >> "$OUT/trace.log" 2>&1
done
grep -v "producer\|mma wait\|mma issue\|dq5" "$OUT/trace.log" | tail -60
echo "== prefill trace gateup M=4096" > "$OUT/trace_prefill.log"
FLUTE_B200_PROFILE=1 timeout 200 python tools/microbench.py --M 4096 --shapes gateup --trace 1 --reps 1 >> "$OUT/trace_prefill.log" 2>&1
grep -v "^     start\|^     setup\|exit" "$OUT/trace_prefill.log" | tail -30
