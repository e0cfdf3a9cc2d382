"""Timing probe (not a test): what bounds the small-K convs (1x1, transposed, 144-channel 3x3)?"""
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
TAPS = {0: 9, 1: 9, 2: 1, 3: 4}
# geom, nb, h, w, ci, co, bias, out_bf16, stats
cases = [(2, 256, 32, 32, 128, 144, 1, 1, 0), (2, 256, 32, 32, 144, 144, 0, 0, 1), (0, 256, 32, 32, 144, 64, 0, 0, 1),
         (3, 256, 16, 16, 256, 128, 1, 1, 0), (0, 256, 32, 32, 144, 128, 0, 0, 1), (2, 256, 32, 32, 128, 128, 0, 1, 0),
         (2, 256, 32, 32, 128, 144, 0, 1, 0), (2, 256, 32, 32, 128, 144, 1, 0, 0)]
for geom, nb, h, w, ci, co, bias, obf, stats in cases:
    x = (torch.rand(nb, h, w, ci, device="cuda") < 0.3).to(torch.bfloat16)
    wgt = (torch.randn(co, TAPS[geom], ci, device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.randn(co, device="cuda") if bias else None
    ho, wo = K.out_hw(geom, h, w)
    out = torch.empty(nb, ho, wo, co, device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
    inb = x.numel() * 2; outb = out.numel() * out.element_size()
    gf = 2.0 * nb * h * w * co * ci * TAPS[geom] / 1e9
    row = []
    for name, knobs in (("default", {}), ("1group", {3: 1}), ("pair", {6: 2}), ("rows-epi", {0: 1}), ("rows+1grp", {0: 1, 3: 1})):
        for k, v in knobs.items(): L.snn_debug_set(k, v)
        try:
            if stats:
                us = t(lambda: K.conv_fprop_partials(geom, x, wgt, co, 4))
            else:
                us = t(lambda: K.conv_fprop(geom, x, wgt, co, bias=b, out=out))
            row.append(f"{name} {us:6.1f}us")
        except Exception as e:
            row.append(f"{name} ERR {str(e)[:40]}")
        for k in knobs: L.snn_debug_set(k, 0)
    print(f"g{geom} {nb}x{h}x{w} {ci}->{co} bias={bias} bf16out={obf} stats={stats}: ideal-mem {(inb + outb) / 6.5e6:5.1f}us ideal-mma {gf / 1.6:5.1f}us | " + " | ".join(row), flush=True)
