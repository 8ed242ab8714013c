#!/bin/bash
# ncu --set full (+ source) of the level-0/1 tensor-core kernels of the second training step:
#   conv3x3_row (forward: enc0.conv_b, enc1.conv_a, enc1.conv_b; later dec2/dec3; dgrads) and the level-0 weight gradients
TAG=${1:-r2l}; OUT=gpurun_out; mkdir -p $OUT
timeout 120 python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_plain.log 2>&1 || exit 1
# one step launches 9 conv3x3_row forwards (3 encoder + 6 decoder incl. none for upconv) and 8-9 dgrads: skip the warm-up step's
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_row" -s 17 -c 17 -f \
  -o $OUT/${TAG}_row python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_ncu1.log 2>&1
ncu -i $OUT/${TAG}_row.ncu-rep --page raw --csv > $OUT/${TAG}_row_raw.csv 2>/dev/null
ls -la $OUT/${TAG}_*
