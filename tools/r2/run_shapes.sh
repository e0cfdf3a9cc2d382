#!/bin/bash
# per-shape in-graph kernel tables for env variants: usage run_shapes.sh <tag> VAR=a VAR=b ...
tag=$1; shift
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; exit 1; }
BA="--no-lif --no-gpu-eager --no-cpu-baseline --no-cfg3 --no-e2e --steps 20 --warmup 3"
for kv in "$@"; do
  env $kv python bench.py $BA > gpurun_out/$tag/${kv}.json 2> gpurun_out/$tag/${kv}.err
  echo "$kv rc=$?"
done
