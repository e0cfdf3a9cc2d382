#!/bin/bash
# light ncu pass (a handful of metrics, few replays) over every conv / wgrad launch of ONE eager step of the final code:
# duration, DRAM bytes, L2->SM bytes, tensor-pipe share.  Usage: run_prof2.sh <tag> [ENV=val ...]
tag=${1:-prof2}; shift
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; exit 1; }
BA="--steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-cfg3 --no-lif --no-gpu-eager --no-profile"
env "$@" python bench.py $BA --dump-trace gpurun_out/$tag/trace.json > gpurun_out/$tag/plain.json 2> gpurun_out/$tag/plain.err; echo "plain rc=$?"
N=$(python -c "import json;t=json.load(open('gpurun_out/$tag/trace.json'));print(sum(1 for n,_ in t if n in ('snn_conv_fprop','snn_conv_fprop_stats','snn_conv_dgrad','snn_conv_wgrad')))")
echo "conv+wgrad launches per step: $N"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
env "$@" timeout -s KILL 1200 ncu --metrics $M --clock-control none -k "regex:^(conv_gemm_kernel|wgrad_gemm_kernel)$" --launch-skip $((3 * N)) --launch-count $N -f -o gpurun_out/$tag/gemm python bench.py $BA > gpurun_out/$tag/gemm.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/$tag/gemm.ncu-rep --page raw --csv > gpurun_out/$tag/gemm.raw.csv 2>/dev/null
rm -f gpurun_out/$tag/gemm.ncu-rep
wc -l gpurun_out/$tag/gemm.raw.csv
