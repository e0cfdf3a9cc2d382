#!/bin/bash
# end-of-round pass: the whole -m gpu suite exactly as the driver runs it, smoke, bench, LIF microbench
mkdir -p gpurun_out/final
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/final/build.log 2>&1 || { echo build failed; exit 1; }
timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/final/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?"; tail -n 4 gpurun_out/final/pytest_gpu.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout -s KILL 600 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
echo "bench rc=$?"; tail -n 2 gpurun_out/final/bench.err
timeout -s KILL 300 python bench.py --microbench lif > gpurun_out/final/lif_microbench.json 2> gpurun_out/final/lif_microbench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final/bench.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"], "\n roofline", d["roofline"], "\n cpu", d["cpu_baseline"], "\n clocks", d["clocks"], d["gpu_launches"])
for k, v in d["kernels"].items():
    print("  %-28s %7.3f ms %5.0f calls %s" % (k, v["ms_per_step"], v["calls_per_step"], {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs")}))
PY
