#!/bin/bash
# NCCL / bucket variants at N GPUs: usage run_ddp_var.sh <tag> <n>
tag=$1; n=$2
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1
run() {
  name=$1; shift
  env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cfg3 --no-e2e --no-profile > gpurun_out/$tag/$name.json 2> gpurun_out/$tag/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$tag/$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "lockstep", d.get("ranks_in_lockstep"))
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/$tag/$name.err").read()[-600:])
PY
}
run default A=1
run simple NCCL_PROTO=Simple
run nvls NCCL_ALGO=NVLS
run bucket128 SNN_BUCKET_MB=128
run bucket512 SNN_BUCKET_MB=512
run ctas8 NCCL_MAX_CTAS=8
run debug NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
grep -i -E "NVLS|algo|proto|channel" gpurun_out/$tag/debug.err | head -20
