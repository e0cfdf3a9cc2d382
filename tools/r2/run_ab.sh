#!/bin/bash
# same-box A/B of env switches: usage run_ab.sh <tag> VAR=a VAR=b ...   (each is one bench run, interleaved twice)
tag=$1; shift
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; exit 1; }
BA="--no-lif --no-gpu-eager --no-cpu-baseline --no-cfg3 --no-profile --no-e2e --steps 40 --warmup 5"
for rep in 1 2; do
  for kv in "$@"; do
    env $kv python bench.py $BA > gpurun_out/$tag/${kv}_$rep.json 2> gpurun_out/$tag/${kv}_$rep.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/$tag/${kv}_$rep.json').read().strip().splitlines()[-1])
print('$kv rep $rep: ms', round(d['ms_per_step'],3), 'value', round(d['value'],1))"
  done
done
