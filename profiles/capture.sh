#!/bin/bash
# End-of-session evidence capture on one B200 (run under gpurun from the repo root):  bash profiles/capture.sh <tag>
# 1. GPU parity suite  2. bench line (with the secondary metrics)  3. per-layer device times
# 4. ncu launch list of two training steps with DRAM bytes + tensor-pipe activity per launch -> per-class traffic of a step
# Each ncu pass runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || exit 1
timeout 300 python profiles/layer_times.py 10 > $OUT/${TAG}_layer_times.csv 2> $OUT/${TAG}_layer_times.err
timeout 120 python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python profiles/step_for_ncu.py 1 1 > $OUT/${TAG}_ncu.log 2>&1
python profiles/step_traffic.py $OUT/${TAG}_launches.csv $OUT/${TAG}_traffic.json \
  "profiles/${TAG}_launches.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active --clock-control none: second of two C2 training steps)" > /dev/null
ls -la $OUT/${TAG}_*
