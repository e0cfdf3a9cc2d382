#!/bin/bash
# usage: bash tools/r2/run_ddp.sh <tag> <ngpus>   (DDP tests at 2 GPUs + torchrun bench at N)
tag=$1; n=$2
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; exit 1; }
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/$tag/smi.txt
if [ "$3" != "notest" ]; then
timeout -s KILL 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q --timeout 600 > gpurun_out/$tag/test_ddp.log 2>&1
echo "ddp tests rc=$?"; tail -n 6 gpurun_out/$tag/test_ddp.log
fi
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/$tag/bench_n$n.json 2> gpurun_out/$tag/bench_n$n.err
echo "bench n$n rc=$?"; tail -n 5 gpurun_out/$tag/bench_n$n.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$tag/bench_n$n.json").read().strip().splitlines()[-1])
    print($n, "value", d["value"], "ms", d["ms_per_step"], "graph", d.get("cuda_graph_active"), "e2e", d["e2e"]["value"] if d.get("e2e") else None, "lockstep", d.get("ranks_in_lockstep"))
    c = d.get("cfg3"); print("cfg3", c and (c["value"], c["ms_per_step"], c["ranks_in_lockstep"]))
    print("other", d.get("other_kernels_ms_per_step"))
except Exception as e:
    print("ERR", e)
PY
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$tag/bench_n$n.json").read().strip().splitlines()[-1])
    print("comm", json.dumps(d.get("comm_timeline"), indent=1))
except Exception as e:
    print("ERR", e)
PY
