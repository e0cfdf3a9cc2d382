#!/bin/bash
# usage: bash tools/r2/run_ncu.sh <tag> <kernel-regex> <launch-skip> <launch-count> [bench args]
tag=$1; rx=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; exit 1; }
BA="--steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-cfg3 --no-lif --no-gpu-eager --no-profile $@"
timeout -s KILL 1500 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip --launch-count $cnt -f -o gpurun_out/$tag/cap python bench.py $BA > gpurun_out/$tag/ncu.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/$tag/ncu.log
ncu -i gpurun_out/$tag/cap.ncu-rep --page raw --csv > gpurun_out/$tag/raw.csv 2>/dev/null
python tools/ncu_brief.py gpurun_out/$tag/raw.csv | cut -c1-260
ls -la gpurun_out/$tag/
