#!/bin/bash
# ncu evidence for the final round-2 kernels (launch list + full captures); sanitizer / host-overhead logs are from r2_profile.sh
set -u
OUT=gpurun_out/r02q
mkdir -p "$OUT"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qgemm_ -s 128 -c 128 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > "$OUT/ncu_launch.log" 2>&1
tail -2 "$OUT/launches.csv" | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:qgemm_decode -s 128 -c 4 \
    -o "$OUT/prof_decode" -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qgemm_prefill -s 2 -c 1 \
    -o "$OUT/prof_prefill" -f python tools/microbench.py --M 4096 --shapes small --reps 1 > "$OUT/ncu_prefill.log" 2>&1
tail -2 "$OUT/ncu_prefill.log"
python tools/host_overhead.py > "$OUT/host_overhead.log" 2>&1
FLUTE_B200_PY_OPS=1 python tools/host_overhead.py >> "$OUT/host_overhead.log" 2>&1
cat "$OUT/host_overhead.log"
