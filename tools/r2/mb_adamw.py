import sys, torch
sys.path.insert(0, ".")
from snn_object_detectionddp_b200 import kernels as K
from snn_object_detectionddp_b200.trainer import one_cycle_table
n = 120_400_000 // 8 * 8
dev = "cuda"
p, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v.abs_()
sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
hp = one_cycle_table(100, 1e-4, 5e-4).float().to(dev)
sumsq = torch.ones(1, device=dev, dtype=torch.float64)
gn = torch.zeros(1, device=dev)
step = torch.zeros(1, device=dev, dtype=torch.int32)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for z in (False, True, False, True):
    ms = t(lambda: K.adamw_step(p, g, m, v, sh, hp, sumsq, gn, step=step, zero_grad=z))
    print("zero_grad", z, "ms", round(ms, 4), "GB/s", round(n * (30 + 4 * z) / ms / 1e6, 1))
ms = t(lambda: g.zero_())
print("fill ms", round(ms, 4))
# with a dirty-L2 producer in front (as in the step: wgrad reduce-adds into g just before)
g2 = torch.randn(n, device=dev)
def seq(z):
    g.add_(g2)            # stand-in for the backward writing g
    K.grad_sumsq(g, sumsq)
    K.adamw_step(p, g, m, v, sh, hp, sumsq, gn, step=step, zero_grad=z)
for z in (False, True):
    print("add+sumsq+adamw zero_grad", z, "ms", round(t(lambda: seq(z)), 4))
