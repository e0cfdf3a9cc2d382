#!/bin/bash
# GPU pass 5: conv (CTA pairs) + neuron (packed-f32x2 backward) tests first with tight timeouts, then the rest, bench, microbench
mkdir -p gpurun_out/r5
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r5/build.log 2>&1 || { echo build failed; tail gpurun_out/r5/build.log; exit 1; }
for f in conv neuron; do
  timeout -s KILL 300 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 120 -x > gpurun_out/r5/$f.log 2>&1
  rc=$?; echo "$f rc=$rc"; grep -E "passed|failed|FAILED|Error|error" gpurun_out/r5/$f.log | tail -n 6
  if [ $rc -ne 0 ]; then tail -n 40 gpurun_out/r5/$f.log; nvidia-smi --query-gpu=name,memory.used --format=csv; exit 1; fi
done
for f in head model train; do
  timeout -s KILL 600 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 300 -x > gpurun_out/r5/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r5/$f.log | tail -n 4
done
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r5/bench.json 2> gpurun_out/r5/bench.err
echo "bench rc=$?"; tail -n 3 gpurun_out/r5/bench.err
timeout -s KILL 300 python bench.py --microbench lif > gpurun_out/r5/lif_microbench.json 2> gpurun_out/r5/lif_microbench.err
echo "lif microbench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r5/bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
    for k, v in d["kernels"].items():
        print("  %-24s %7.3f ms %5.0f calls %s" % (k, v["ms_per_step"], v["calls_per_step"], {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs")}))
    for r in (d.get("kernels_by_shape") or [])[:40]:
        print("   ", r)
except Exception as e:
    print("bench parse failed", e)
try:
    d = json.loads(open("gpurun_out/r5/lif_microbench.json").read().strip().splitlines()[-1])
    for r in d["rows"]:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("T", "C", "HW", "fwd_frac", "bwd_frac", "bwd_reduce_gbs", "bwd_dx_gbs", "fwd_gbs")})
except Exception as e:
    print("microbench parse failed", e)
PY
