#!/bin/bash
# End-of-session evidence capture on one B200 (run under gpurun from the repo root):  bash profiles/capture.sh <tag>
# 1. GPU parity suite  2. bench line  3. per-layer device times  4. ncu launch list of one training step
# 5. ncu --set full of every tensor-core launch of one training step (the .ncu-rep is summarised by ncu_summary.py)
# Each ncu pass runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
TAG=${1:-r1k}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || exit 1
timeout 300 python profiles/layer_times.py 10 > $OUT/${TAG}_layer_times.csv 2> $OUT/${TAG}_layer_times.err
timeout 120 python profiles/step_for_ncu.py 2 1 > $OUT/${TAG}_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 140 --csv \
  --log-file $OUT/${TAG}_launches.csv python profiles/step_for_ncu.py 2 1 > $OUT/${TAG}_ncu.log 2>&1
if [ "$2" == "full" ]; then
  # backward half of the second step: 21 weight gradients + 20 dgrads (the first 63 + 21 tensor-core launches are step 1 and
  # step 2's forward).  The report is summarised on the box (raw page as CSV) and deleted: gpurun returns at most 64 MiB.
  timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"wgrad3x3_halo|conv3x3_halo|conv3x3_row" -s 84 -c 42 -f -o /tmp/${TAG}_prof \
    python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_ncu_full.log 2>&1
  ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_raw.csv 2>> $OUT/${TAG}_ncu_full.log
fi
ls -la $OUT/${TAG}_*
