#!/bin/bash
# ncu pass: launch list of a short bench run + full captures of the top kernels (1 GPU)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
for k in conv_gemm_kernel:60 wgrad_gemm_kernel:50 bn_act_bwd_kernel:20 bn_act_fwd_kernel:22 bn_stats_kernel:20; do
  name=${k%%:*}; cnt=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -c $cnt -o gpurun_out/prof_$name -f $CMD > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"
done
ls -la gpurun_out | tail -n 15
tail -n 3 gpurun_out/plain.log
