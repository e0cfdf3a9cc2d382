#!/bin/bash
# quick pass: neuron + model tests, bench, LIF microbench
mkdir -p gpurun_out/r7
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r7/build.log 2>&1 || { echo build failed; exit 1; }
for f in neuron model; do
  timeout -s KILL 400 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 200 -x > gpurun_out/r7/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r7/$f.log | tail -n 3
done
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r7/bench.json 2> gpurun_out/r7/bench.err
echo "bench rc=$?"; tail -n 3 gpurun_out/r7/bench.err
timeout -s KILL 300 python bench.py --microbench lif > gpurun_out/r7/lif_microbench.json 2> gpurun_out/r7/lif_microbench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r7/bench.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d["kernels"].items():
    if "bn_act" in k: print("  %-28s %7.3f ms %5.0f calls %s" % (k, v["ms_per_step"], v["calls_per_step"], {a: round(b, 1) for a, b in v.items() if a in ("tflops", "gbs")}))
for r in d["kernels_by_shape"]:
    if "bwd2" in r["kernel"]: print("   ", r)
d = json.loads(open("gpurun_out/r7/lif_microbench.json").read().strip().splitlines()[-1])
for r in d["rows"]:
    if r["C"] in (128, 512) and r["HW"] in (64, 16): print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("T", "C", "HW", "fwd_frac", "bwd_frac", "bwd_reduce_gbs", "bwd_dx_gbs")})
PY
