"""-m gpu: the path bench.py times, at the widths and batch shapes it times them, on ONE GPU.

* `Trainer.train_step_graphed` (CUDA-graph replay, what the bench measures) == `Trainer.train_step` (eager launches) from
  identical state: bit-equal forward (loss items), gradients / Adam moments within the fp32-atomics reordering noise.
* one full-width training step at BASELINE.json configs[1] (T=4, B=64, 256x256) and configs[2] (T=8, B=16, 512x512) against
  the oracle port of the reference step (oracle/model_oracle.py on CUDA, bf16-operand contract): loss items, global
  gradient norm, per-layer spike flip report (LIF) -- reference train.py:58-80 semantics.
* the reference's native frame size 480x640 (its own smoke shape, model.py:213-219; DSEC frames are never resized,
  dataset.py:151): output maps (64,80),(32,40),(16,20) through the bilinear skip-resize branch (model.py:43-44).

Stated tolerances are in each test.
"""
import pytest
import torch

from oracle import model_oracle as MO
from oracle import snn_oracle as O
from tests.gpu_util import rel_err, setup_exact

pytestmark = pytest.mark.gpu
DEV = "cuda"
HYP = {"box": 7.5, "cls": 1.0, "dfl": 2.5, "reg_max": 16}


def _product(neuron, seed=0, state=None):
    from snn_object_detectionddp_b200.model import YOLOTemporalUNet
    from snn_object_detectionddp_b200.weight_initialization import initialize_model
    torch.manual_seed(seed)
    net = YOLOTemporalUNet(num_classes=8, hyp=HYP, neuron=neuron)
    initialize_model(net)
    if state is not None:
        res = net.load_state_dict(state, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    return net.to(DEV)


def _oracle(neuron, seed=0):
    torch.manual_seed(seed)
    orc = MO.OracleYOLOTemporalUNet(num_classes=8, hyp=HYP, neuron=neuron, emulate_bf16=True)
    MO.initialize_model_oracle(orc)
    return orc


def _share_frozen_extractor(orc, net):
    """Both sides read the SAME frozen features.  The extractor is a frozen, gradient-free stand-in (the reference's
    pretrained YOLO11m cannot exist offline) whose own parity is tested separately
    (test_feature_extractor_standin_matches_oracle, 4e-3); its bf16 output roundings occasionally differ from the fp32
    restatement by one bf16 ulp, and a spiking net amplifies every such ulp into a cascade of flips -- that would test the
    stand-in, not the path.  Product features: bf16 NHWC -> fp32 NCHW, exactly the values the product's U-Net consumes."""
    orc.feature_extractor = net.feature_extractor


def _sync_state(src, dst):
    """Copy trainer `src`'s whole training state into `dst` IN PLACE (a captured graph holds the addresses)."""
    for name in ("flat_p", "flat_m", "flat_v", "shadow", "flat_g"):
        getattr(dst.store, name).copy_(getattr(src.store, name))
    for a, b in zip(src.model.buffers(), dst.model.buffers()):
        b.copy_(a)
    dst._step_dev.copy_(src._step_dev)
    dst.step_idx = src.step_idx


def test_graphed_step_equals_eager_step_full_width():
    """Same state in, same batch: the replayed graph and the eager launch sequence run the same kernels.
    Forward has no atomics -> loss items bit-equal.  Backward sums fp32 partials with atomics / TMA reduce-add in a
    non-fixed order and rounds dy to bf16: gradient norm and Adam first moment agree to 2e-3 (measured 1e-7 / 3e-4,
    printed; a wrong/missing kernel in the captured graph would show as O(1))."""
    setup_exact()
    from snn_object_detectionddp_b200.data import synthetic_batch
    from snn_object_detectionddp_b200.trainer import Trainer
    B, T, HW = 16, 4, 256
    a = Trainer(_product("lif", seed=3), total_steps=50, device=DEV)
    b = Trainer(_product("lif", seed=3), total_steps=50, device=DEV)
    worst = 0.0
    for step in range(6):
        frames, labels = synthetic_batch(B, T, HW, HW, seed=200 + step)
        frames = frames.to(DEV)
        batch = {"padded": tuple(t.to(DEV) for t in a.prepare_batch(labels, B, max_boxes=8)["padded"])}
        _sync_state(a, b)
        _, it_a = a.train_step(frames, batch)
        it_a = it_a.clone()
        _, it_b = b.train_step_graphed(frames, batch)
        it_b = it_b.clone()
        torch.cuda.synchronize()
        assert torch.equal(it_a, it_b), (step, it_a, it_b)
        e_m = rel_err(b.store.flat_m, a.store.flat_m)          # Adam first moment: linear in the (clipped) gradient
        e_p = rel_err(b.store.flat_p, a.store.flat_p)
        e_n = abs(float(a.grad_norm) - float(b.grad_norm)) / float(a.grad_norm)
        print(f"step {step}: graphed={b._graph is not None} loss {it_a.tolist()} exp_avg rel {e_m:.2e} params rel {e_p:.2e} norm rel {e_n:.2e}")
        worst = max(worst, e_m, e_n)
        assert float(a.grad_norm) > 0 and float(a.store.flat_m.abs().max()) > 0
        # the optimizer pass leaves the gradient buffer zeroed (optimizer.zero_grad() of train.py:61 folded in)
        assert float(a.store.flat_g.abs().max()) == 0 and float(b.store.flat_g.abs().max()) == 0
    assert b._graph is not None and not b._graph_failed, "the step was never captured"
    assert worst < 2e-3, worst


def _flip_report(name, s_prod, s_orc, u_orc, theta=1.0):
    """Spikes [T,B,C,H,W].  flips = disagreeing spikes; near = flips whose ORACLE membrane is within 1e-5 of threshold
    (north_star's window); carried = flips of a neuron that already had a near-threshold flip at an EARLIER timestep (the
    LIF membrane is state: one flipped reset changes that neuron's next membrane by up to beta*theta, teacher-forcing the
    layer INPUT cannot undo it); unexplained = the rest, with the largest |u - theta| among them."""
    flips = s_prod != s_orc
    dist = (u_orc - theta).abs()
    near = flips & (dist < 1e-5)
    seeded = torch.zeros_like(near)
    seeded[1:] = near.cummax(0).values[:-1]           # a near-threshold flip happened strictly before t in this neuron
    carried = flips & ~near & seeded
    unexplained = flips & ~near & ~seeded
    md = float(dist[unexplained].max()) if bool(unexplained.any()) else 0.0
    return dict(layer=name, n=s_orc.numel(), flips=int(flips.sum()), near=int(near.sum()), carried=int(carried.sum()),
                unexplained=int(unexplained.sum()), max_dist_unexplained=md, far=int((flips & ~near).sum()), rate=float(s_orc.float().mean()))


@pytest.mark.parametrize("neuron,B,T,HW", [("silu", 64, 4, 256), ("lif", 64, 4, 256), ("lif", 16, 8, 512)],
                         ids=["cfg2-silu", "cfg2-lif", "cfg3-lif"])
def test_full_width_training_step_vs_oracle(neuron, B, T, HW):
    """BASELINE.json configs[1] / configs[2] at full width, one GPU, two optimizer steps.

    silu (the reference's own network): loss items 2e-3, gradient norm 5e-3 (measured 2.4e-4 / 5e-4 at B=64: large
    BatchNorm groups average the bf16 rounding noise of the 23 layers).
    lif (build-defined): flip-rate protocol -- teacher-forced per layer every spike agrees with the oracle except neurons
    whose oracle membrane is within 1e-5 of threshold; end to end the near-threshold flips cascade (reported); loss items 0.1."""
    setup_exact()
    from snn_object_detectionddp_b200.trainer import Trainer
    orc = _oracle(neuron, seed=5)
    net = _product(neuron, state=orc.state_dict())
    orc = orc.to(DEV).train()
    _share_frozen_extractor(orc, net)
    frames, labels = MO.synthetic_batch(B, T, HW, HW, seed=21)
    frames, labels = frames.to(DEV), labels.to(DEV)
    loss_fn, opt, sched = MO.make_reference_trainer(orc, total_steps=20)
    tr = Trainer(net, total_steps=20, device=DEV)
    batch = {"batch_idx": labels[:, 0], "cls": labels[:, 1], "bboxes": labels[:, 2:]}
    tol_loss, tol_gn = (2e-3, 5e-3) if neuron == "silu" else (0.1, 0.25)
    if neuron == "lif":
        # flip-rate protocol at full width.  (1) teacher-forced: every ConvBlock of the product is fed the ORACLE's input
        # of that layer -> spikes must agree except neurons whose oracle membrane is within 1e-5 of threshold;
        # (2) end to end: the first layer obeys the same bound; behind it every near-threshold flip changes ~9*Cout
        # downstream membranes by one weight and cascades (a spiking net is chaotic) -- rates are printed, never hidden.
        import snn_object_detectionddp_b200.model as M
        from snn_object_detectionddp_b200.params import store_for
        names = {m: n for n, m in orc.temporal_unet.named_modules() if isinstance(m, O.OracleConvBlock)}
        rec_s, rec_u, rec_x = {}, {}, {}

        def hook(m, inp, outp):
            rec_x.setdefault(names[m], []).append(inp[0].detach().to(torch.bfloat16))     # bf16-representable by construction
            rec_s.setdefault(names[m], []).append(outp[0].detach().to(torch.bool))
            rec_u.setdefault(names[m], []).append(m.last_u)

        hs = [m.register_forward_hook(hook) for m in names]
        with torch.no_grad():
            hid = None
            for t in range(T):
                _, hid = orc(frames[:, t], hid)
            rec = {}
            net.train()
            net.forward_sequence(frames, record=rec)
        for h in hs:
            h.remove()
        e2e, forced = [], []
        blocks = dict(net.temporal_unet.named_modules())
        st = store_for(net, DEV)
        st.refresh_operands()
        for name, s_list in rec_s.items():
            s_o, u_o = torch.stack(s_list), torch.stack(rec_u[name])                 # [T,B,C,H,W]
            sp = rec[name]
            e2e.append(_flip_report(name, sp.reshape(T, B, *sp.shape[1:]).permute(0, 1, 4, 2, 3) > 0.5, s_o, u_o))
            x_o = torch.cat(rec_x[name], 0)
            assert torch.equal(x_o.float().to(torch.bfloat16), x_o)
            with torch.no_grad():
                out, _ = blocks[name].forward_seq(M.RunCtx(st, T), x_o.permute(0, 2, 3, 1).contiguous())
            forced.append(_flip_report(name, out.reshape(T, B, *out.shape[1:]).permute(0, 1, 4, 2, 3) > 0.5, s_o, u_o))
            del x_o, out
        del rec_s, rec_u, rec_x, rec
        print("\nLIF flip report, teacher-forced per layer:", *forced, sep="\n  ")
        print("LIF flip report, end to end (first forward):", *e2e, sep="\n  ")
        assert len(forced) == 16
        # Measured on B200: 0-33 flips per layer out of 4-67 M neuron-steps (<= 1.5e-6): every one is either inside north_star's
        # 1e-5 window or carried by the same neuron's membrane from such a flip at an earlier timestep; anything else would be
        # "unexplained" (allowed only within 1e-4 of threshold = fp32 accumulation-order noise of a K <= 9216 conv after BN).
        # (flip COUNT: ~6e-6 * n membranes lie inside the 2e-5-wide window; which of them land on the other side depends on the
        # fp32 summation order of the conv, which differs between torch and the kernel's stencil-column order: <= 3e-6 * n + 8)
        assert all(r["flips"] <= 3e-6 * r["n"] + 8 for r in forced), forced
        assert all(r["max_dist_unexplained"] < 1e-4 for r in forced), [r for r in forced if r["max_dist_unexplained"] >= 1e-4]
        assert sum(r["unexplained"] for r in forced) <= 16, forced                 # measured: 3 (cfg2) / 8 (cfg3) of 2.7e8 / 5.4e8, all < 3e-5
        assert all(0.02 < r["rate"] < 0.7 for r in forced), forced
        assert e2e[0]["layer"] == "enc1" and e2e[0]["max_dist_unexplained"] < 1e-4 and e2e[0]["flips"] <= 3e-6 * e2e[0]["n"] + 8, e2e[0]
        # (the forwards above advanced the BatchNorm running statistics of both sides; train mode does not read them)
    for step in range(2):
        _, it_o, gn_o = MO.reference_train_step(orc, loss_fn, opt, sched, frames, labels)
        _, it_p = tr.train_step(frames, batch)
        print(f"{neuron} B={B} T={T} {HW}x{HW} step {step}: oracle {it_o.tolist()} gn {float(gn_o):.4f} | product {it_p.tolist()} gn {float(tr.grad_norm):.4f}")
        assert torch.allclose(it_p, it_o, rtol=tol_loss, atol=1e-3), (step, it_p, it_o)
        assert abs(float(tr.grad_norm) - float(gn_o)) < tol_gn * float(gn_o), (step, float(tr.grad_norm), float(gn_o))


@pytest.mark.parametrize("neuron", ["silu", "lif"])
def test_native_480x640_frames_through_the_skip_resize_branch(neuron):
    """The reference's own smoke shape (model.py:213-219: 2 x 3 x 480 x 640): P3/P4/P5 = 60x80 / 30x40 / 15x20, the
    stride-2 conv on the odd 15-row level gives 8x10, up-sampling gives 16x20 / 32x40 / 64x80 and the three skips are
    bilinearly resized (model.py:43-44).  Output maps must be (64,80),(32,40),(16,20) and agree with the
    oracle port (bf16 operand contract): silu 2e-2 per map; lif: shapes, finiteness, loss within 0.25 (spike drift)."""
    setup_exact()
    from snn_object_detectionddp_b200.loss import v8DetectionLoss
    from snn_object_detectionddp_b200.params import store_for
    orc = _oracle(neuron, seed=9)
    net = _product(neuron, state=orc.state_dict())
    orc = orc.to(DEV).train()
    _share_frozen_extractor(orc, net)
    net.train()
    B, T = 2, 2
    g = torch.Generator().manual_seed(33)
    frames = torch.rand(B, T, 3, 480, 640, generator=g).to(DEV)
    _, labels = MO.synthetic_batch(B, 1, 64, 64, seed=34)
    labels = labels.to(DEV)
    batch = {"batch_idx": labels[:, 0], "cls": labels[:, 1], "bboxes": labels[:, 2:]}
    # drop-in per-frame loop (train.py:62-66)
    hid_o = hid_p = None
    for t in range(T):
        preds_o, hid_o = orc(frames[:, t], hid_o)
        preds_p, hid_p = net(frames[:, t], hid_p)
    assert [tuple(p.shape) for p in preds_p] == [(B, 72, 64, 80), (B, 72, 32, 40), (B, 72, 16, 20)]
    assert [tuple(p.shape) for p in preds_o] == [tuple(p.shape) for p in preds_p]
    assert tuple(hid_p[0].shape) == (B, 1024, 8, 10)
    errs = [rel_err(a, b) for a, b in zip(preds_p, preds_o)]
    print(f"\n480x640 {neuron}: per-map rel err vs oracle {errs}")
    assert all(torch.isfinite(p).all() for p in preds_p)
    if neuron == "silu":
        assert max(errs) < 2e-2, errs
    l_o, it_o = MO.D.OracleV8DetectionLoss(orc)(preds_o, batch)
    l_p, it_p = v8DetectionLoss(net)(preds_p, batch)
    st = store_for(net, DEV)
    st.zero_grad()
    l_p.sum().backward()
    l_o.sum().backward()
    print(f"   loss oracle {it_o.tolist()} product {it_p.tolist()}")
    assert torch.allclose(it_p, it_o, rtol=3e-2 if neuron == "silu" else 0.25, atol=1e-3)
    gn_p = float(st.flat_g.double().norm())
    gn_o = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in orc.parameters() if p.grad is not None)))
    print(f"   grad norm oracle {gn_o:.4f} product {gn_p:.4f}")
    assert gn_p > 0 and abs(gn_p - gn_o) < (6e-2 if neuron == "silu" else 0.4) * gn_o
    # fused sequence path == the per-frame loop on the same shapes (deterministic forward)
    net2 = _product(neuron, state=orc.state_dict())
    net2.train()
    with torch.no_grad():
        det, _ = net2.forward_sequence(frames)
    for a, b in zip(det.maps_nchw(), preds_p):
        assert torch.equal(a, b.detach())
