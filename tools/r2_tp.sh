#!/bin/bash
# Tensor-parallel validation on N GPUs of one box: fused-exchange test + the bench at --gpus N.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_tp.sh 2'
set -u
N=${1:-2}
OUT=gpurun_out/r02tp$N
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "tp_fused" -s > "$OUT/pytest_tp.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_tp.log"; tail -15 "$OUT/pytest_tp.log"
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 50 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?"; tail -12 "$OUT/bench.err"; cat "$OUT/bench.json"
