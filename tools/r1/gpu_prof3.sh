#!/bin/bash
# ncu pass v2 (after the CTA-pair / TMA-store / packed-f32x2 rewrites): launch list of eager steps + full captures -> CSV
O=gpurun_out/prof3
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-profile --no-graph"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -n 20 $O/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name regex skip count keep_rep
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $O/$1 $CMD > $O/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  if [ "$5" != "keep" ]; then rm -f $O/$1.ncu-rep; fi
}
cap conv_fprop conv_gemm_kernel 6 14 keep
cap conv_dgrad conv_gemm_kernel 60 12
cap wgrad wgrad_gemm_kernel 20 8
cap lif_bwd bn_act_bwd2_kernel 40 6 keep
cap lif_fwd bn_act_fwd_kernel 3 5
cap bn_partials bn_stats_from_partials_kernel 0 3
du -sh $O; ls -la $O
