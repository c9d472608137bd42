mkdir -p gpurun_out/${1:-r01p}
export FLUTE_B200_PROFILE=1
python tools/microbench.py --M 4096 --shapes small --trace 1 --reps 2 2>&1 | tee gpurun_out/${1:-r01p}/trace_small_M4096.log
python tools/microbench.py --M 1 --shapes gateup --trace 1 --reps 3 2>&1 | tee gpurun_out/${1:-r01p}/trace_gateup_M1.log
