#!/bin/bash
# 2-GPU pass: NCCL DDP tests + torchrun bench at N=2 (and N=1 on the same box for the ratio)
mkdir -p gpurun_out/ddp
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/ddp/build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/ddp/smi.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q --timeout 600 -x > gpurun_out/ddp/test_ddp.log 2>&1
echo "ddp tests rc=$?"; tail -n 15 gpurun_out/ddp/test_ddp.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ddp/bench_n1.json 2> gpurun_out/ddp/bench_n1.err
echo "bench n1 rc=$?"
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/ddp/bench_n2.json 2> gpurun_out/ddp/bench_n2.err
echo "bench n2 rc=$?"; tail -n 5 gpurun_out/ddp/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/ddp/bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d.get("cuda_graph_active"), d["e2e"]["value"] if d.get("e2e") else None)
    except Exception as e:
        print(n, "ERR", e)
PY
