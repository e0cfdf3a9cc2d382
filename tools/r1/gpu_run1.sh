#!/bin/bash
# first GPU pass: neuron kernels, then conv kernels (each group in its own process, bounded)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout -s KILL 600 python -m pytest tests/test_gpu_neuron.py -m gpu -q --timeout 120 > gpurun_out/neuron.log 2>&1
echo "neuron rc=$?" >> gpurun_out/summary.txt
for grp in fprop dgrad wgrad weight_prep; do
  timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 90 -k "$grp" > gpurun_out/conv_$grp.log 2>&1
  echo "conv_$grp rc=$?" >> gpurun_out/summary.txt
done
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/neuron.log
tail -n 15 gpurun_out/conv_fprop.log
