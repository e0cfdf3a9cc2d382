#!/bin/bash
# GPU pass 4: full -m gpu suite (bounded groups), bench (CUDA graph), LIF microbench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in conv neuron head model train; do
  timeout -s KILL 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 300 -x > gpurun_out/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/summary.txt
done
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/summary.txt
timeout -s KILL 300 python bench.py --microbench lif > gpurun_out/lif_microbench.json 2> gpurun_out/lif_microbench.err
echo "lif microbench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in conv neuron head model train; do grep -E "passed|failed|FAILED|Error" gpurun_out/$f.log | tail -n 6; done
tail -n 5 gpurun_out/bench.err
