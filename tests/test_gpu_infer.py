"""-m gpu: inference path (BASELINE.json configs[4]) -- snn_nms vs the restated ultralytics / torchvision NMS (bit-exact on
identical predictions: indices, order and values), and the eval-mode pipeline (T-frame unroll -> decode -> NMS) vs the
oracle pipeline.  Head / decode / NMS are PARITY UNPINNED (ultralytics is un-vendored; oracle/detect_oracle.py is the
specification)."""
import pytest
import torch

from oracle import detect_oracle as D
from oracle import model_oracle as MO
from tests.gpu_util import rel_err, setup_exact

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _synthetic_pred(B, nc, A, seed, n_obj=12, hw=256.0, dup_scores=False):
    """[B, 4+nc, A]: clusters of overlapping boxes around a few objects + background clutter."""
    g = torch.Generator().manual_seed(seed)
    pred = torch.zeros(B, 4 + nc, A)
    for b in range(B):
        centers = torch.rand(n_obj, 2, generator=g) * hw
        sizes = 10 + torch.rand(n_obj, 2, generator=g) * 60
        which = torch.randint(0, n_obj, (A,), generator=g)
        jitter = torch.randn(A, 4, generator=g) * 3.0
        pred[b, 0:2] = (centers[which] + jitter[:, :2]).t()
        pred[b, 2:4] = (sizes[which] + jitter[:, 2:].abs()).t()
        sc = torch.rand(A, nc, generator=g) ** 4                      # mostly low scores, a few high
        if dup_scores:
            sc = (sc * 16).round() / 16                               # many exact ties
        pred[b, 4:] = sc.t()
    if B > 1:
        pred[-1, 4:] = 0.0                                            # an image without detections
    return pred.to(DEV)


CASES = [
    # B, nc, A, conf, iou, multi_label, agnostic, max_det, dup
    (2, 8, 1344, 0.3, 0.45, True, False, 300, False),     # visualize.py:73-78
    (3, 8, 1344, 0.001, 0.6, False, False, 300, False),   # eval_2.py:108 (every anchor is a candidate)
    (2, 8, 5376, 0.05, 0.45, True, False, 300, False),    # 512x512 maps, ~40 % of 43 k pairs pass -> global-memory sort
    (2, 8, 1344, 0.3, 0.45, True, True, 300, False),      # class-agnostic
    (2, 8, 1344, 0.25, 0.5, True, False, 300, True),      # exact score ties -> enumeration-order tie break
    (1, 1, 6720, 0.1, 0.7, True, False, 50, False),       # single class (multi_label is ignored), max_det binds
    (4, 3, 336, 0.5, 0.45, False, False, 300, False),
]


@pytest.mark.parametrize("B,nc,A,conf,iou,multi,agn,max_det,dup", CASES)
def test_nms_bit_exact_vs_oracle(B, nc, A, conf, iou, multi, agn, max_det, dup):
    from snn_object_detectionddp_b200.nms import non_max_suppression
    pred = _synthetic_pred(B, nc, A, seed=B * 1000 + A + nc, dup_scores=dup)
    ours, idxs = non_max_suppression(pred, conf, iou, multi_label=multi, agnostic=agn, max_det=max_det, return_idxs=True)
    ref = D.non_max_suppression(pred.cpu(), conf, iou, multi_label=multi, agnostic=agn, max_det=max_det)
    assert len(ours) == len(ref) == B
    total = 0
    for b in range(B):
        o, r = ours[b].cpu(), ref[b]
        assert o.shape == r.shape, (b, o.shape, r.shape)
        assert torch.equal(o, r), (b, (o != r).nonzero()[:5])
        total += o.shape[0]
        # the reported enumeration index points at the row's (anchor, class)
        if o.shape[0]:
            e = idxs[b].cpu()
            a_idx, c_idx = (e // nc, e % nc) if (multi and nc > 1) else (e, o[:, 5].long())
            assert torch.equal(c_idx.float(), o[:, 5])
            assert torch.equal(pred[b, 4 + c_idx.to(DEV), a_idx.to(DEV)].cpu(), o[:, 4])
    assert total > 0
    if B > 1:
        assert ours[-1].shape[0] == 0


def _models(seed=21):
    from snn_object_detectionddp_b200.model import TemporalUNet, YOLOTemporalUNet
    hyp = {"box": 7.5, "cls": 1.0, "dfl": 2.5, "reg_max": 16}
    widths = (64, 128, 256, 512)
    torch.manual_seed(seed)
    orc = MO.OracleYOLOTemporalUNet(num_classes=8, hyp=hyp, neuron="lif", emulate_bf16=True, widths=widths)
    MO.initialize_model_oracle(orc)
    with torch.no_grad():                       # a head that actually fires: positive class-logit bias
        for seq in orc.detection_head.cv3:
            seq[-1].bias.fill_(-0.8)
    net = YOLOTemporalUNet(num_classes=8, hyp=hyp, neuron="lif")
    net.temporal_unet = TemporalUNet([144, 144, 144], neuron="lif", widths=widths)
    res = net.load_state_dict(orc.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return orc.to(DEV).eval(), net.to(DEV).eval()


@pytest.mark.parametrize("B", [1, 32])
def test_inference_pipeline_vs_oracle(B):
    """Batch-1 and batch-32, T=4 windows (visualize.py:62-78 / eval_2.py:96-112): eval-mode unroll -> decode -> NMS.
    The product's NMS on the product's own prediction equals the oracle NMS on that same tensor bit for bit; against the
    oracle PIPELINE (fp32 torch convs on bf16-rounded operands) predictions agree to the bf16-conv tolerance and the
    detections are matched (same class, box corners within 2 px) for the large majority: with a random-init head many
    boxes overlap at IoU close to the threshold, so the greedy choice is sensitive to the bf16-level differences."""
    setup_exact()
    from snn_object_detectionddp_b200.infer import decode_last_step, detect_sequence
    from snn_object_detectionddp_b200.nms import non_max_suppression
    orc, net = _models()
    T, HW = 4, 128
    frames, _ = MO.synthetic_batch(B, T, HW, HW, seed=31)
    frames = frames.to(DEV)
    pred, _ = decode_last_step(net, frames)
    assert pred.shape == (B, 12, 336)
    conf, iou = 0.3, 0.45
    ours, idxs = non_max_suppression(pred, conf, iou, multi_label=True, return_idxs=True)
    same_input = D.non_max_suppression(pred.cpu(), conf, iou, multi_label=True)
    for o, r in zip(ours, same_input):
        assert torch.equal(o.cpu(), r)
    via_api = detect_sequence(net, frames, conf_thres=conf, iou_thres=iou, multi_label=True)
    for o, r in zip(ours, via_api):
        assert torch.equal(o, r)
    # oracle pipeline: per-frame loop with state carry (visualize.py:66-71), eval mode
    with torch.no_grad():
        hid = None
        for t in range(T):
            (o_pred, _), hid = orc(frames[:, t], hid)
    assert rel_err(pred, o_pred) < 2e-2
    ref = D.non_max_suppression(o_pred.cpu(), conf, iou, multi_label=True)
    n_ours, n_ref, matched = 0, 0, 0
    for b in range(B):
        o, r = ours[b].cpu(), ref[b]
        n_ours += o.shape[0]; n_ref += r.shape[0]
        for row in o:
            hit = ((r[:, 5] == row[5]) & ((r[:, :4] - row[:4]).abs().max(1).values < 2.0)) if r.shape[0] else torch.zeros(0, dtype=torch.bool)
            matched += int(hit.any())
    print(f"B={B}: detections ours {n_ours} oracle {n_ref} matched {matched}")
    assert n_ref > 0 and n_ours > 0
    assert abs(n_ours - n_ref) <= 0.05 * n_ref + 2 and matched >= 0.7 * max(n_ours, n_ref)


def test_streaming_detector_carries_state():
    """StreamingDetector(T=1 chunks) == one windowed call over the same frames (state threaded through `hidden`)."""
    setup_exact()
    from snn_object_detectionddp_b200.infer import StreamingDetector, decode_last_step
    _, net = _models(seed=22)
    frames, _ = MO.synthetic_batch(2, 3, 128, 128, seed=32)
    frames = frames.to(DEV)
    pred_win, _ = decode_last_step(net, frames)
    hid = None
    for t in range(3):
        pred_t, hid = decode_last_step(net, frames[:, t:t + 1], hid, return_state=True)
    assert torch.equal(pred_win, pred_t)
    det = StreamingDetector(net, conf_thres=0.3)
    for t in range(3):
        out = det(frames[:, t])
    assert len(out) == 2


@pytest.mark.parametrize("B", [1, 32])
def test_eval_spikes_teacher_forced_bit_compare(B):
    """configs[4]: batch-1 / batch-32 streaming windows, T=4, eval mode (BatchNorm running statistics).  Every spiking
    ConvBlock of the product is fed the ORACLE's input of that layer (teacher forcing, so one layer's flips cannot
    cascade): spikes must agree exactly, except neurons whose oracle membrane lies within 1e-5 of threshold -- those are
    counted and reported as the flip rate (SURVEY.md 7.2 protocol)."""
    setup_exact()
    from oracle import snn_oracle as O
    from snn_object_detectionddp_b200.model import RunCtx
    from snn_object_detectionddp_b200.params import store_for
    orc, net = _models(seed=23)
    with torch.no_grad():                       # non-trivial running statistics
        for m in orc.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 2.0)
    net.load_state_dict(orc.state_dict())
    orc.eval(); net.eval()
    T, HW = 4, 128
    frames, _ = MO.synthetic_batch(B, T, HW, HW, seed=33)
    frames = frames.to(DEV)
    names = {m: n for n, m in orc.temporal_unet.named_modules() if isinstance(m, O.OracleConvBlock)}
    rec_x, rec_s, rec_u = {}, {}, {}

    def hook(m, inp, outp):
        rec_x.setdefault(names[m], []).append(inp[0].detach())
        rec_s.setdefault(names[m], []).append(outp[0].detach())
        rec_u.setdefault(names[m], []).append(m.last_u)

    hs = [m.register_forward_hook(hook) for m in names]
    with torch.no_grad():
        hid = None
        for t in range(T):
            _, hid = orc(frames[:, t], hid)
    for h in hs:
        h.remove()
    st = store_for(net, DEV)
    st.refresh_operands()
    blocks = dict(net.temporal_unet.named_modules())
    report, n_tot, flips, far = [], 0, 0, 0
    with torch.no_grad():
        for name, xs in rec_x.items():
            x = torch.cat(xs, 0).permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()          # folded [T*B,h,w,C]
            out, _ = blocks[name].forward_seq(RunCtx(st, T), x)
            s_p = out.reshape(T, B, *out.shape[1:]).permute(0, 1, 4, 2, 3).float()
            s_o, u_o = torch.stack(rec_s[name]), torch.stack(rec_u[name])
            d = s_p != s_o
            near = (u_o - 1.0).abs() < 1e-5
            report.append((name, int(d.sum()), int((d & ~near).sum()), float(s_o.mean())))
            n_tot += s_o.numel(); flips += int(d.sum()); far += int((d & ~near).sum())
    print(f"B={B}: eval spike flip report (layer, flips, far-from-threshold flips, spike rate):", *report, sep="\n  ")
    print(f"B={B}: flip rate {flips}/{n_tot} = {flips / n_tot:.2e}")
    assert len(report) == 16 and far == 0 and flips <= max(4, int(2e-6 * n_tot)), report


def test_graphed_window_detector_equals_eager():
    """CUDA-graph replay of the eval window (unroll -> decode -> NMS) == the eager calls, on changing inputs."""
    setup_exact()
    from snn_object_detectionddp_b200.infer import GraphedWindowDetector, detect_sequence
    _, net = _models(seed=24)
    det = GraphedWindowDetector(net, conf_thres=0.3)
    for seed in (41, 42, 43):
        frames, _ = MO.synthetic_batch(2, 3, 128, 128, seed=seed)
        frames = frames.to(DEV)
        rows_g, kept_g, counts_g = (t.clone() for t in det(frames))
        rows_e, kept_e, counts_e = detect_sequence(net, frames, conf_thres=0.3, padded=True)
        assert torch.equal(counts_g, counts_e)
        for b in range(2):
            n = int(counts_e[b])
            assert torch.equal(rows_g[b, :n], rows_e[b, :n]) and torch.equal(kept_g[b, :n], kept_e[b, :n])


def test_nms_kernel_reproduces_torchvision_fixture(golden_dir):
    """snn_nms vs the committed fixture recorded from the real torchvision.ops.nms (tests/golden/make_golden_nms.py):
    rows, order and values bit-exact -- no oracle code involved at comparison time."""
    import os
    from snn_object_detectionddp_b200.nms import non_max_suppression
    fx = torch.load(os.path.join(golden_dir, "nms_golden.pt"), weights_only=False)
    for case in fx["cases"]:
        kw = dict(case["kwargs"])
        ours = non_max_suppression(case["pred"].to(DEV), kw.pop("conf_thres"), kw.pop("iou_thres"), max_det=300, **kw)
        for a, b in zip(ours, case["rows"]):
            assert torch.equal(a.cpu(), b), case["name"]
