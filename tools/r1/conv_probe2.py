"""Timing probe (not a test): small-K convs, where the per-tile skeleton (epilogue + barriers) dominates."""
import sys, torch
sys.path.insert(0, ".")
from snn_object_detectionddp_b200 import kernels as K, _lib
L = _lib.lib()
def t(fn, reps=30):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
cases = [(2, 256, 32, 32, 128, 144, torch.bfloat16), (2, 256, 32, 32, 144, 144, torch.float32), (3, 256, 16, 16, 256, 128, torch.bfloat16),
         (0, 256, 32, 32, 144, 64, torch.float32), (0, 256, 32, 32, 128, 128, torch.float32)]
for geom, nb, h, w, ci, co, dt in cases:
    taps = {0: 9, 2: 1, 3: 4}[geom]
    x = (torch.rand(nb, h, w, ci, device="cuda") < 0.3).to(torch.bfloat16)
    wgt = (torch.randn(co, taps, ci, device="cuda") * 0.02).to(torch.bfloat16)
    ho, wo = K.out_hw(geom, h, w)
    out = torch.empty(nb, ho, wo, co, device="cuda", dtype=dt)
    mb = (x.numel() * 2 + out.numel() * out.element_size()) / 1e6
    for single, legacy in ((0, 0), (1, 0), (0, 1), (1, 1)):
        row = []
        for dbg in (0, 3):
            L.snn_debug_set(6, single); L.snn_debug_set(0, legacy); L.snn_debug_set(7, dbg)
            us = t(lambda: K.conv_fprop(geom, x, wgt, co, out=out))
            row.append("dbg%d %6.1fus %5.0fGB/s" % (dbg, us, mb / us * 1e3 / 1e3))
        print(f"g{geom} {nb}x{h}x{w} {ci}->{co} {str(dt)[6:]} single={single} rows_epi={legacy}: " + " | ".join(row), flush=True)
L.snn_debug_set(6, 0); L.snn_debug_set(7, 0); L.snn_debug_set(0, 0)
