#!/bin/bash
# One gpurun call: GPU parity tests, per-shape microbench, the judged bench, the ncu launch list and one
# full ncu capture of the qgemm kernel.  Everything lands in gpurun_out/ (copy what matters to profiles/).
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
fi

timeout 300 python tools/microbench.py --M 1,16,512,4096 --shapes llama8b --json "$OUT/microbench_llama8b.json" > "$OUT/microbench_llama8b.log" 2>&1
cat "$OUT/microbench_llama8b.log"

timeout 400 python bench.py --steps 30 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"

if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list: the two eager warm-up tokens of bench.py are 256 direct launches
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qgemm_ -s 128 -c 128 --csv \
      --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_launch.log" 2>&1
  # full capture: gate_up (28672x4096), o (4096x4096) decode launches of the second eager token
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:qgemm_decode -s 129 -c 2 \
      -o "$OUT/prof_decode" -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
  tail -2 "$OUT/ncu_full.log"
  if [ "${NCU_PREFILL:-1}" = "1" ]; then
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:qgemm_prefill -s 2 -c 1 \
        -o "$OUT/prof_prefill" -f python tools/microbench.py --M 4096 --shapes small --reps 1 > "$OUT/ncu_prefill.log" 2>&1
    tail -2 "$OUT/ncu_prefill.log"
  fi
fi
ls -la "$OUT"
