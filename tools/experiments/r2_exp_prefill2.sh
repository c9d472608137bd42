#!/bin/bash
# Prefill partial-tile hand-over (coalesced slots, paired loads): parity, then Stream-K vs whole tiles at mid-size M,
# and the 16-accumulator decode variant against the general kernel at M = 8 / 16.
set -u
OUT=gpurun_out/r02p2
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_qgemm_gpu.py -x -q -m gpu -k "prefill or tile_p or config2 or schedules_agree or numerics_bound" > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest.log"; tail -4 "$OUT/pytest.log" | cut -c1-200
for sk in -1 0 1; do
  echo "== force-streamk $sk"
  timeout 300 python tools/microbench.py --M 128,256,512,1024,2048,4096 --shapes llama8b --force-streamk $sk 2>&1 | grep "N="
done > "$OUT/prefill_streamk.log" 2>&1
cat "$OUT/prefill_streamk.log"
for v in -1 2; do
  echo "== variant $v"
  timeout 300 python tools/microbench.py --M 5,8,16 --shapes llama8b --variant $v 2>&1 | grep "N="
done > "$OUT/m16_variant.log" 2>&1
cat "$OUT/m16_variant.log"
