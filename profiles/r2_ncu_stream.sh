#!/bin/bash
# ncu --set full of the CUDA-core / streaming kernels of one training step (second step; the first warms up):
#   first layer (forward, weight gradient), head, BatchNorm forward (level-0 dropout / pool, level 1),
#   BatchNorm backward (first launches of the step = level 0).  bash profiles/r2_ncu_stream.sh <tag>
TAG=${1:-r2h}; OUT=gpurun_out; mkdir -p $OUT
timeout 120 python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"c1_fwd8|c1_8|head_kernel" -s 3 -c 3 -f \
  -o $OUT/${TAG}_c1_head python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"bn_apply_kernel" -s 17 -c 3 -f \
  -o $OUT/${TAG}_bn_fwd python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"bn_bwd|relu_bwd" -s 38 -c 7 -f \
  -o $OUT/${TAG}_bn_bwd python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_ncu3.log 2>&1
for f in c1_head bn_fwd bn_bwd; do
  ncu -i $OUT/${TAG}_$f.ncu-rep --page raw --csv > $OUT/${TAG}_${f}_raw.csv 2>/dev/null
done
ls -la $OUT/${TAG}_*
