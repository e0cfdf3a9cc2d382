#!/bin/bash
# GPU pass 3: model + train tests, then first bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in model train; do
  timeout -s KILL 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 300 -s > gpurun_out/$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/summary.txt
done
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/model.log | tail -n 20
grep -E "passed|failed|FAILED|Error" gpurun_out/train.log | tail -n 30
tail -n 5 gpurun_out/bench.err
cat gpurun_out/bench.json
