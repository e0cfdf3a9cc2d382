#!/bin/bash
# ncu pass (compact): launch list of one eager step + small full captures, exported to CSV on the box.
mkdir -p gpurun_out/prof
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in neuron model; do
  timeout -s KILL 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 300 -x > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/$f.log | tail -n 4
done
timeout -s KILL 300 python bench.py --microbench lif > gpurun_out/lif_microbench.json 2> gpurun_out/lif_microbench.err
echo "lif microbench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-profile --no-graph"
$CMD > gpurun_out/prof/plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/prof/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/prof/launches.csv $CMD > gpurun_out/prof/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof/$1 -f $CMD > gpurun_out/prof/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/prof/$1.ncu-rep --page raw --csv > gpurun_out/prof/$1.raw.csv 2>/dev/null
}
cap conv_fprop conv_gemm_kernel 6 12
cap conv_dgrad conv_gemm_kernel 70 10
cap wgrad wgrad_gemm_kernel 20 8
cap lif_bwd bn_act_bwd2_kernel 40 6
cap lif_fwd bn_act_fwd_kernel 3 5
cap bn_stats bn_stats_kernel 0 3
du -sh gpurun_out/prof; ls -la gpurun_out/prof
# keep the bundle under the 64 MiB copy-back limit: drop the biggest reports first (their CSVs stay)
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do
  big=$(ls -S gpurun_out/prof/*.ncu-rep 2>/dev/null | head -n 1); [ -z "$big" ] && break; echo "dropping $big"; rm -f "$big"
done
du -sh gpurun_out
