"""static vs dynamic tile scheduling on bench shapes (single GPU, timing only)"""
import sys, torch
sys.path.insert(0, ".")
from snn_object_detectionddp_b200 import kernels as K, _lib
L = _lib.lib()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (nb, h, w, ci, co) in [(256, 32, 32, 128, 128), (256, 16, 16, 256, 256), (256, 8, 8, 512, 512), (256, 4, 4, 1024, 1024)]:
    x = (torch.rand(nb, h, w, ci, device="cuda") < 0.3).to(torch.bfloat16)
    wgt = (torch.randn(co, 9, ci, device="cuda") * 0.02).to(torch.bfloat16)
    dy = torch.randn(nb, h, w, co, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(co, 9, ci, device="cuda")
    row = []
    for dyn in (0, 1):
        L.snn_set_tile_scheduling(dyn)
        row.append("dyn%d fprop %6.1f dgrad %6.1f wgrad %6.1f us" % (dyn, t(lambda: K.conv_fprop_partials(0, x, wgt, co, 4)),
                   t(lambda: K.conv_dgrad(0, dy, wgt, (h, w), ci)), t(lambda: K.conv_wgrad(0, x, dy, dw))))
    L.snn_set_tile_scheduling(0)
    print(f"{nb}x{h}x{w} {ci}->{co}: " + " | ".join(row), flush=True)
