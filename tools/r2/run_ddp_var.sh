#!/bin/bash
# tile-scheduling / bucket variants at N GPUs: usage run_ddp_var.sh <tag> <n>
tag=$1; n=$2
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cfg3 --no-e2e --no-profile > gpurun_out/$tag/$name.json 2> gpurun_out/$tag/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$tag/$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "lockstep", d.get("ranks_in_lockstep"))
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/$tag/$name.err").read()[-600:])
PY
}
run dyn1_b32 SNN_DYNAMIC_TILES=1
run dyn0_b32 SNN_DYNAMIC_TILES=0
run dyn1_b128 SNN_DYNAMIC_TILES=1 SNN_BUCKET_MB=128
