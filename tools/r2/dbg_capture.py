"""Which part of the training step breaks CUDA-graph capture?  Each stage in its own process (a failed capture poisons the context)."""
import subprocess
import sys

STAGES = ["E_unet_bwd_only", "full"]

if len(sys.argv) == 1:
    for s in STAGES:
        r = subprocess.run([sys.executable, __file__, s], capture_output=True, text=True)
        tail = (r.stdout + r.stderr).strip().splitlines()[-40:]
        print(f"== {s}: rc={r.returncode}")
        for l in tail:
            print("   ", l[:300])
    sys.exit(0)

import torch

sys.path.insert(0, ".")
from snn_object_detectionddp_b200.data import synthetic_batch
from snn_object_detectionddp_b200.model import YOLOTemporalUNet
from snn_object_detectionddp_b200.trainer import Trainer
from snn_object_detectionddp_b200.weight_initialization import initialize_model

stage = sys.argv[1]
HYP = {"box": 7.5, "cls": 1.0, "dfl": 2.5, "reg_max": 16}
torch.manual_seed(0)
net = YOLOTemporalUNet(num_classes=8, hyp=HYP, neuron="lif")
initialize_model(net)
tr = Trainer(net, total_steps=50, device="cuda")
B, T, HW = 4, 2, 128
frames, labels = synthetic_batch(B, T, HW, HW, seed=1)
frames = frames.cuda()
batch = {"padded": tuple(t.cuda() for t in tr.prepare_batch(labels, B, max_boxes=8)["padded"])}
if stage.endswith("noout"):
    from snn_object_detectionddp_b200 import ops as _ops
    _orig_fwd = _ops.ConvBiasFn.forward

    def _fwd(ctx, x0, weight, bias, cfg):
        cfg = dict(cfg)
        cfg.pop("out", None)
        return _orig_fwd(ctx, x0, weight, bias, cfg)

    _ops.ConvBiasFn.forward = staticmethod(_fwd)
if stage.endswith("nowarm"):
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        for _ in range(2):
            tr.train_step(frames, batch)
    torch.cuda.synchronize()
else:
    for _ in range(2):
        tr.train_step(frames, batch)
torch.cuda.synchronize()


def body():
    net.train()
    tr.store.zero_grad()
    det, _ = net.forward_sequence(frames)
    if stage == "full":
        loss, items = tr.loss_fn(det, batch)
        torch.autograd.backward((loss,), (tr._ones3,))
        tr.optimizer_step()
        return items
    if stage.startswith("A"):
        loss, items = tr.loss_fn(det.maps_nchw(), batch)
        torch.autograd.backward((loss,), (tr._ones3,))
        return items
    if stage.startswith("B"):
        loss, items = tr.loss_fn(det, batch)
        loss.sum().backward()
        return items
    if stage.startswith("C"):
        loss, items = tr.loss_fn(det, batch)
        gs = torch.autograd.grad((loss,), tuple(det.box) + tuple(det.cls), (tr._ones3,))
        return gs[0].float()
    if stage.startswith("D"):
        gb = [torch.ones_like(b, dtype=torch.bfloat16) for b in det.box]
        gc = [torch.ones_like(c, dtype=torch.bfloat16) for c in det.cls]
        hd = [p for p in net.detection_head.parameters() if p.requires_grad]
        torch.autograd.backward(tuple(det.box) + tuple(det.cls), tuple(gb) + tuple(gc), inputs=hd)
        return det.flat_box
    if stage.startswith("E"):
        gb = [torch.ones_like(b, dtype=torch.bfloat16) for b in det.box]
        gc = [torch.ones_like(c, dtype=torch.bfloat16) for c in det.cls]
        torch.autograd.backward(tuple(det.box) + tuple(det.cls), tuple(gb) + tuple(gc))
        return det.flat_box


# ---- trace the capture status around every libsnnb200 call ----
import ctypes
from snn_object_detectionddp_b200 import _lib, kernels as K, ops, loss as L_, head as H_, nms, params
rt = ctypes.CDLL("libcudart.so.12") if True else None
_orig_call = _lib.call
state = {"bad": False, "n": 0}


def cap_status():
    st = ctypes.c_int(-1)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value


def traced_call(name, *args, work=None):
    before = cap_status()
    _orig_call(name, *args, work=work)
    after = cap_status()
    state["n"] += 1
    if not state["bad"] and (before[1] == 2 or after[1] == 2 or before[0] != 0 or after[0] != 0):
        state["bad"] = True
        import traceback
        print(f"!! capture invalidated around call #{state['n']} {name}: before={before} after={after}")
        traceback.print_stack(limit=8)


for mod in (_lib, K, nms):
    if hasattr(mod, "call"):
        mod.call = traced_call

g = torch.cuda.CUDAGraph()
kw = {"capture_error_mode": "relaxed"} if stage.endswith("relaxed") else {}
with torch.cuda.graph(g, **kw):
    out = body()
g.replay()
torch.cuda.synchronize()
print("captured + replayed OK", out.flatten()[:3].tolist())
