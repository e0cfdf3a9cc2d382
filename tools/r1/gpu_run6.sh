#!/bin/bash
# GPU pass 6: whole -m gpu suite in bounded groups, bench, LIF microbench
mkdir -p gpurun_out/r6
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r6/build.log 2>&1 || { echo build failed; tail gpurun_out/r6/build.log; exit 1; }
for f in conv neuron infer head model train; do
  timeout -s KILL 600 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 200 -x > gpurun_out/r6/$f.log 2>&1
  rc=$?; echo "$f rc=$rc"; grep -E "passed|failed|FAILED|Error" gpurun_out/r6/$f.log | tail -n 4
  if [ $rc -ne 0 ] && [ "$f" = "conv" ]; then tail -n 30 gpurun_out/r6/$f.log; exit 1; fi
done
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r6/bench.json 2> gpurun_out/r6/bench.err
echo "bench rc=$?"; tail -n 3 gpurun_out/r6/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r6/bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
    for k, v in d["kernels"].items():
        print("  %-28s %7.3f ms %5.0f calls %s" % (k, v["ms_per_step"], v["calls_per_step"], {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs")}))
    for r in (d.get("kernels_by_shape") or [])[:14]:
        print("   ", r)
except Exception as e:
    print("bench parse failed", e)
PY
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
