#!/bin/bash
# ncu --set full of EVERY tensor-core kernel of the second training step (forward convs, dgrads, weight gradients):
#   bash profiles/r2_ncu_tc.sh <tag>   -> gpurun_out/<tag>_tc_raw.csv
TAG=${1:-r2zf}; OUT=gpurun_out; mkdir -p $OUT
timeout 120 python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_plain.log 2>&1 || exit 1
# count the tensor-core launches of one step first (cheap pass), then skip the warm-up step's
N=$(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"conv3x3_row|conv3x3_halo|conv3x3_tc|wgrad3x3" \
    python profiles/step_for_ncu.py 1 1 2>/dev/null | grep -c "gpu__time_duration.sum")
HALF=$((N / 2))
echo "tensor-core launches: $N in two steps, capturing the last $HALF" > $OUT/${TAG}_tc.log
timeout 1500 ncu --set full --clock-control none -k regex:"conv3x3_row|conv3x3_halo|conv3x3_tc|wgrad3x3" -s $HALF -c $HALF -f \
  -o $OUT/${TAG}_tc python profiles/step_for_ncu.py 1 1 >> $OUT/${TAG}_tc.log 2>&1
ncu -i $OUT/${TAG}_tc.ncu-rep --page raw --csv > $OUT/${TAG}_tc_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_tc.ncu-rep
ls -la $OUT/${TAG}_tc*
