#!/bin/bash
# quick GPU check of the decode kernel: parity tests, then microbench + timeline
OUT=gpurun_out/${1:-r01c}; mkdir -p $OUT
[ -x tools/mma_probe.bin ] && [ "${PROBE:-0}" = "1" ] && timeout 60 ./tools/mma_probe.bin | tee $OUT/mma_probe.log
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python tools/microbench.py --M ${MS:-1,16} --shapes llama8b > $OUT/mb_decode.log 2>&1; cat $OUT/mb_decode.log
timeout 120 python tools/microbench.py --M 1 --shapes gateup --trace 1 --reps 3 > $OUT/trace_gateup.log 2>&1; cat $OUT/trace_gateup.log
timeout 120 python tools/microbench.py --M 1 --shapes small --trace 1 --reps 3 > $OUT/trace_small.log 2>&1; cat $OUT/trace_small.log
