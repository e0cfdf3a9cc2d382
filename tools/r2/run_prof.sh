#!/bin/bash
# round-2 evidence: launch list of the bench command + full ncu captures of the conv / LIF kernels of ONE eager training step
tag=${1:-prof}
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; exit 1; }
BA="--steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-cfg3 --no-lif --no-gpu-eager --no-profile"
python bench.py $BA --dump-trace gpurun_out/$tag/trace.json > gpurun_out/$tag/plain.json 2> gpurun_out/$tag/plain.err; echo "plain rc=$?"
# (1) launch list (durations only) of the graphed bench command: kernels inside the replayed graph are profiled per node
timeout -s KILL 900 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/$tag/launches_graph.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-cfg3 --no-lif --no-gpu-eager --no-profile > gpurun_out/$tag/launches_graph.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/$tag/launches_graph.csv)"
# (2) full captures.  One eager step = the launches after the warm-up step's; skip counts = launches of the family in the two
# untimed steps (dump-trace runs two extra eager steps only in the plain run above, not here): warm-up (1 step).
NCONV=$(python -c "import json;t=json.load(open('gpurun_out/$tag/trace.json'));print(sum(1 for n,_ in t if n in ('snn_conv_fprop','snn_conv_fprop_stats','snn_conv_dgrad')))")
NWG=$(python -c "import json;t=json.load(open('gpurun_out/$tag/trace.json'));print(sum(1 for n,_ in t if n=='snn_conv_wgrad'))")
NLIF=$(python -c "import json;t=json.load(open('gpurun_out/$tag/trace.json'));print(sum(1 for n,_ in t if n in ('snn_bn_act_fwd','snn_bn_act_bwd2')))")
echo "per step: conv_gemm $NCONV wgrad $NWG lif $NLIF"
cap() { # name regex skip count
  timeout -s KILL 1500 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$tag/$1 python bench.py $BA > gpurun_out/$tag/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/$tag/$1.ncu-rep --page raw --csv > gpurun_out/$tag/$1.raw.csv 2>/dev/null
  rm -f gpurun_out/$tag/$1.ncu-rep
}
# bench.py runs max(warmup, 3) = 3 untimed eager steps before the timed one: skip their launches
cap conv "^conv_gemm_kernel$" $((3 * NCONV)) $NCONV
cap wgrad "^wgrad_gemm_kernel$" $((3 * NWG)) $NWG
cap lif "bn_act_fwd_kernel|bn_act_bwd2_kernel|silu_t1_bwd2_kernel" $((3 * NLIF)) $NLIF
ls -la gpurun_out/$tag
