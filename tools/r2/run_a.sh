#!/bin/bash
# round 2, GPU pass A: new parity tests (bench shapes, multi-tile, real-reference blocks, 480x640, graphed==eager), full suite, bench
mkdir -p gpurun_out/a
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/a/build.log; exit 1; }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a/smi.txt
for f in test_gpu_blocks_reference test_gpu_bench_shapes test_gpu_fullwidth; do
  timeout -s KILL 1200 python -m pytest tests/$f.py -q -m gpu -s --timeout 900 > gpurun_out/a/$f.log 2>&1
  echo "$f rc=$?"; tail -n 3 gpurun_out/a/$f.log
done
timeout -s KILL 1200 python -m pytest tests/ -q -m gpu --timeout 900 --deselect tests/test_gpu_blocks_reference.py --deselect tests/test_gpu_bench_shapes.py --deselect tests/test_gpu_fullwidth.py > gpurun_out/a/pytest_rest.log 2>&1
echo "rest rc=$?"; tail -n 5 gpurun_out/a/pytest_rest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout -s KILL 600 python bench.py --no-cpu-baseline > gpurun_out/a/bench.json 2> gpurun_out/a/bench.err
echo "bench rc=$?"; tail -n 2 gpurun_out/a/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/a/bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
