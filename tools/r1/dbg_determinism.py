"""Debug aid: run the same training step twice and report the first libsnnb200 call whose outputs differ."""
import sys, torch
sys.path.insert(0, ".")
from tests.test_gpu_train import _models, MO, DEV
from tests.gpu_util import setup_exact
from snn_object_detectionddp_b200 import kernels as K, _lib
from snn_object_detectionddp_b200.params import store_for
from snn_object_detectionddp_b200.loss import v8DetectionLoss
setup_exact()
log = None
orig_call = _lib.call
import snn_object_detectionddp_b200.kernels as KM

def wrap(name, fn):
    def f(*a, **k):
        out = fn(*a, **k)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        sig = []
        for o in outs:
            if isinstance(o, torch.Tensor):
                d = o.detach().double()
                sig.append((float(d.sum()), float(d.abs().sum())))
        # in-place accumulators passed as args
        for o in a:
            if isinstance(o, torch.Tensor) and o.dtype in (torch.float32, torch.float64) and name in ("conv_wgrad", "colsum_accumulate", "dw3x3_wgrad"):
                d = o.detach().double()
                sig.append((float(d.sum()), float(d.abs().sum())))
        log.append((name, tuple(tuple(o.shape) for o in outs if isinstance(o, torch.Tensor)), tuple(sig)))
        return out
    return f

for n in dir(KM):
    fn = getattr(KM, n)
    if callable(fn) and getattr(fn, "__module__", "") == KM.__name__ and not n.startswith("_") and n not in ("out_hw",):
        setattr(KM, n, wrap(n, fn))

logs = []
for run in range(2):
    log = []
    _, net = _models("lif", seed=7)
    net.skip_dead_backward = False
    net.train()
    B, T, HW = 2, 3, 128
    frames, labels = MO.synthetic_batch(B, T, HW, HW, seed=13)
    frames, labels = frames.to(DEV), labels.to(DEV)
    st = store_for(net, DEV)
    st.zero_grad()
    det, _ = net.forward_sequence(frames)
    loss, items = v8DetectionLoss(net)(det, {"batch_idx": labels[:, 0], "cls": labels[:, 1], "bboxes": labels[:, 2:]})
    loss.sum().backward()
    torch.cuda.synchronize()
    logs.append(log)
a, b = logs
print(len(a), len(b))
nd = 0
for i, (x, y) in enumerate(zip(a, b)):
    if x != y:
        print("DIFF at call", i, x[0], x[1])
        for s1, s2 in zip(x[2], y[2]):
            print("    ", s1, s2)
        nd += 1
        if nd >= 6:
            break
print("first calls:", [x[0] for x in a[:5]], "... total diffs shown", nd)
