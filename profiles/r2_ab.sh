#!/bin/bash
# A/B of an environment toggle on one box: bash profiles/r2_ab.sh <tag> <ENVVAR> [steps]
# bench line (no extras) and per-layer profile for ENVVAR unset and ENVVAR=1
TAG=$1; VAR=$2; STEPS=${3:-20}
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1; do
  if [ $v == 1 ]; then export $VAR=1; else unset $VAR; fi
  timeout 300 python bench.py --no-extra --steps $STEPS > $OUT/${TAG}_${VAR}${v}_bench.json 2> $OUT/${TAG}_${VAR}${v}_bench.err
  timeout 200 python profiles/layer_times.py 10 > $OUT/${TAG}_${VAR}${v}_layers.csv 2>> $OUT/${TAG}_${VAR}${v}_bench.err
done
python - <<PY
import json
for v in (0,1):
    try:
        d=json.load(open('$OUT/${TAG}_${VAR}%d_bench.json'%v))
        print('$VAR=%d'%v, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:x['ms_per_step'] for k,x in d['roofline']['kernels'].items()})
    except Exception as e: print('fail',v,e)
PY
