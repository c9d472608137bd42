#!/bin/bash
# General kernel, 3-bit decode (M = 1, fp16): CTAs per launch swept per Llama-3-8B shape.
set -u
OUT=gpurun_out/r02g3
mkdir -p "$OUT"
timeout 600 python tools/microbench.py --bits 3 --dtype fp16 --M 1 --reps 10 --force-grid 0,144,128,112,96,72,64,48,32 --shapes llama8b 2>&1 | grep "N=" > "$OUT/grid_sweep_w3.log"
cut -c1-60,118- "$OUT/grid_sweep_w3.log"
