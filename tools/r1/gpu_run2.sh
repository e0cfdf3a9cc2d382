#!/bin/bash
# GPU pass 2: full -m gpu suite in groups (each group its own bounded process)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in neuron head model; do
  timeout -s KILL 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 300 -s > gpurun_out/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
tail -n 60 gpurun_out/model.log
tail -n 15 gpurun_out/head.log
tail -n 5 gpurun_out/neuron.log
