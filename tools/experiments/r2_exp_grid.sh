#!/bin/bash
# Decode kernel, M = 1: CTAs per launch swept per shape -- the four Llama-3-8B shapes and their tensor-parallel column shards.
set -u
OUT=gpurun_out/r02g2
mkdir -p "$OUT"
G=0,136,128,112,96,80,64,48,32
timeout 600 python tools/microbench.py --M 1 --reps 10 --force-grid $G \
  --shapes 6144x4096,4096x4096,28672x4096,4096x14336,3072x4096,2048x4096,14336x4096,2048x14336,1536x4096,1024x4096,7168x4096,1024x14336,768x4096,512x4096,3584x4096,512x14336 \
  2>&1 | grep "N=" > "$OUT/grid_sweep.log"
cut -c1-60,118- "$OUT/grid_sweep.log"
