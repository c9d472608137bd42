mkdir -p gpurun_out/${1:-r01h}
export FLUTE_B200_PROFILE=1
for s in gateup small; do
python tools/microbench.py --M 1 --shapes $s --trace 1 --reps 3 2>&1 | tee gpurun_out/${1:-r01h}/trace_${s}_M1.log
done
for ab in 7; do python tools/microbench.py --M 1 --shapes gateup --trace 1 --reps 3 --ablate $ab 2>&1 | tee gpurun_out/${1:-r01h}/trace_gateup_abl$ab.log; done
