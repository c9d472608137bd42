#!/bin/bash
# Round-2 evidence run: ncu launch list + full captures (decode, prefill), compute-sanitizer, host overhead of the bindings.
set -u
OUT=gpurun_out/r02p
mkdir -p "$OUT"
python tools/host_overhead.py > "$OUT/host_overhead.log" 2>&1
FLUTE_B200_PY_OPS=1 python tools/host_overhead.py >> "$OUT/host_overhead.log" 2>&1
cat "$OUT/host_overhead.log"
echo "== M=16: general kernel (default) vs decode kernel MC=16 (variant 2)" > "$OUT/microbench_M16.log"
timeout 200 python tools/microbench.py --M 16 --shapes llama8b >> "$OUT/microbench_M16.log" 2>&1
timeout 200 python tools/microbench.py --M 16 --shapes llama8b --variant 2 >> "$OUT/microbench_M16.log" 2>&1
cat "$OUT/microbench_M16.log"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qgemm_ -s 128 -c 128 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > "$OUT/ncu_launch.log" 2>&1
tail -2 "$OUT/launches.csv" | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:qgemm_decode -s 128 -c 4 \
    -o "$OUT/prof_decode" -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qgemm_prefill -s 2 -c 1 \
    -o "$OUT/prof_prefill" -f python tools/microbench.py --M 4096 --shapes small --reps 1 > "$OUT/ncu_prefill.log" 2>&1
tail -2 "$OUT/ncu_prefill.log"
# compute-sanitizer: race and barrier checks on a slice of the GPU suite that touches all three kernels
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python -m pytest tests -m gpu -x -q \
      -k "edge_shapes or decode_kernel_identity or prefill_kernel_identity or golden_vectors" > "$OUT/sanitizer_$tool.log" 2>&1
  echo "$tool exit $?" >> "$OUT/sanitizer_$tool.log"; tail -6 "$OUT/sanitizer_$tool.log" | cut -c1-200
done
ls -la "$OUT"
