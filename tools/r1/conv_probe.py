"""Timing probe for the conv GEMM main loop (not a test): which resource bounds it?
dbg bit 0 = no TMA loads, bit 1 = no MMAs; flag 6 = 1 forces the single-CTA kernel; flag 1 caps the stage count."""
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
shapes = [(0, 256, 16, 16, 256, 256), (0, 256, 4, 4, 1024, 4096), (0, 256, 32, 32, 128, 128), (0, 256, 8, 8, 512, 512)]
for geom, nb, h, w, ci, co in shapes:
    x = (torch.rand(nb, h, w, ci, device="cuda") < 0.3).to(torch.bfloat16)
    wgt = (torch.randn(co, 9, ci, device="cuda") * 0.02).to(torch.bfloat16)
    out = torch.empty(nb, h, w, co, device="cuda")
    gf = 2.0 * nb * h * w * co * ci * 9 / 1e9
    for single, legacy in ((0, 0), (1, 0), (0, 1)):
        L.snn_debug_set(0, legacy)
        for stages in (0,):
            row = []
            for dbg in ((0, 3) if (h, ci) != (32, 128) else (0,)):
                L.snn_debug_set(6, single); L.snn_debug_set(7, dbg); L.snn_debug_set(1, stages)
                us = t(lambda: K.conv_fprop(geom, x, wgt, co, out=out))
                row.append("dbg%d %7.1fus %6.0fTF" % (dbg, us, gf / us * 1e3))
            print(f"{nb}x{h}x{w} {ci}->{co} single={single} rows_epilogue={legacy} stages={stages or 'max'}: " + " | ".join(row), flush=True)
L.snn_debug_set(6, 0); L.snn_debug_set(7, 0); L.snn_debug_set(1, 0); L.snn_debug_set(0, 0)
