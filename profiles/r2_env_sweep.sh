#!/bin/bash
# bash profiles/r2_env_sweep.sh <tag> "VAR=VAL ..." "VAR=VAL ..." ...   (one quick_step.py run per quoted environment; "" = default)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for cfg in "$@"; do
  env $cfg timeout 200 python profiles/quick_step.py "${cfg:-default}" 2>&1 | tail -1 | tee -a $OUT/${TAG}_sweep.jsonl
done
