#!/bin/bash
# T == 16 chunked LIF backward: neuron tests, then the configs[3] sweep
tag=${1:-lif}
bash tools/r2/run_tests.sh $tag tests/test_gpu_neuron.py
timeout -s KILL 600 python bench.py --microbench lif > gpurun_out/$tag/lif.json 2> gpurun_out/$tag/lif.err; echo "lif rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/$tag/lif.json").read().strip().splitlines()[-1])
l=d.get("lif_microbench", d)
for r in l["rows"]:
    print(r["T"], r["C"], r["HW"], "fwd %.2f bwd %.2f  (reduce %.0f dx %.0f GB/s)" % (r["fwd_frac"], r["bwd_frac"], r["bwd_reduce_gbs"], r["bwd_dx_gbs"]))
print(l.get("clocks"))
PY
