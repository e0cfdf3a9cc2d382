"""Timing probe (not a test): BASELINE.json configs[4] inference path -- eval-mode T=4 window -> decode -> NMS, batch 1 and 32."""
import json, sys, torch
sys.path.insert(0, ".")
from snn_object_detectionddp_b200.model import YOLOTemporalUNet
from snn_object_detectionddp_b200.weight_initialization import initialize_model
from snn_object_detectionddp_b200.infer import GraphedWindowDetector, detect_sequence
HYP = {"box": 7.5, "cls": 1.0, "dfl": 2.5, "reg_max": 16}
torch.manual_seed(42)
model = YOLOTemporalUNet(num_classes=8, hyp=HYP, neuron="lif")
initialize_model(model)
model = model.cuda().eval()
rows = []
for B in (1, 32):
    frames = torch.randint(0, 256, (B, 4, 3, 256, 256), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        out = detect_sequence(model, frames, conf_thres=0.3, iou_thres=0.45, multi_label=True, padded=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        out = detect_sequence(model, frames, conf_thres=0.3, iou_thres=0.45, multi_label=True, padded=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gd = GraphedWindowDetector(model, conf_thres=0.3, iou_thres=0.45, multi_label=True)
    for _ in range(3):
        gd(frames)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(n):
        gd(frames)
    g1.record(); torch.cuda.synchronize()
    gms = g0.elapsed_time(g1) / n
    rows.append({"batch": B, "graph_ms_per_window": gms, "graph_windows_per_s": B / gms * 1e3, "T": 4, "hw": 256, "ms_per_window": ms, "windows_per_s": B / ms * 1e3, "frames_per_s": 4 * B / ms * 1e3,
                 "detections_last": int(out[2].sum())})
print(json.dumps({"probe": "inference window (eval unroll + decode + NMS, eager launches, no host sync)", "rows": rows}))
