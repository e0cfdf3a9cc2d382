"""TEST INFRASTRUCTURE ONLY -- CPU/any-device restatement of the hot path in plain PyTorch.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file.  The product (``snn_object_detectionddp_b200``) never does.

Two classes of behaviour are restated here:

* **Reference-pinned** -- module topology and math of ``/root/reference/model.py``:
  ``ConvBlock`` (model.py:9-18), ``DownBlock`` (model.py:20-30), ``UpBlock`` (model.py:32-48),
  ``ConvLSTM2d`` (model.py:50-71), ``TemporalUNet`` (model.py:100-146), weight init
  (weight_initialization.py:8-56) and the T-step unroll of train.py:62-66.  With
  ``neuron='silu'`` these classes are pinned against the *real* reference classes by
  ``tests/golden/make_golden.py`` (fixtures committed) and ``tests/test_oracle_vs_reference.py``.

* **Build-defined (PARITY UNPINNED)** -- the reference contains no LIF neuron, surrogate gradient
  or spike code.  ``neuron='lif'`` swaps ``SiLU`` for the LIF neuron of SURVEY.md section 7.2:
      u[t] = beta*v[t-1] + x[t];  s[t] = (u[t] >= theta);  v[t] = u[t]*(1-s[t])   (hard reset)
      ds/du ~= (alpha/2) / (1 + (pi*alpha*(u-theta)/2)^2)   (ATan surrogate, reset not detached)
  There is nothing in the reference to pin this against; it is the specification the CUDA
  kernels are tested against.

All arithmetic is fp32 (or fp64 when the caller casts the module).  ``emulate_bf16=True`` rounds
conv *operands* (activations and weights) to bf16 before an fp32-accumulated conv, which is the
numeric contract of the tcgen05 kernels.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

LIF_DEFAULTS = dict(beta=0.5, v_th=1.0, alpha=2.0)


# ----------------------------------------------------------------------------------------------
# build-defined LIF neuron
# ----------------------------------------------------------------------------------------------
class _ATanSpike(torch.autograd.Function):
    """Heaviside forward, ATan surrogate backward."""

    @staticmethod
    def forward(ctx, u, theta, alpha):
        ctx.save_for_backward(u)
        ctx.theta, ctx.alpha = theta, alpha
        return (u >= theta).to(u.dtype)

    @staticmethod
    def backward(ctx, gs):
        (u,) = ctx.saved_tensors
        a = ctx.alpha
        z = (math.pi * a / 2.0) * (u - ctx.theta)
        return gs * (a / 2.0) / (1.0 + z * z), None, None


def lif_step(x, v_prev, beta, theta, alpha):
    """One LIF timestep. Evaluation order is part of the spec: u = (beta*v) + x with two roundings."""
    u = beta * v_prev + x
    s = _ATanSpike.apply(u, theta, alpha)
    v = u * (1.0 - s)
    return s, v, u


def lif_sequence(x_seq, v0=None, beta=0.5, theta=1.0, alpha=2.0):
    """x_seq: [T, ...] input currents. Returns spikes [T,...], final membrane, pre-spike u [T,...]."""
    v = torch.zeros_like(x_seq[0]) if v0 is None else v0
    ss, us = [], []
    for t in range(x_seq.shape[0]):
        s, v, u = lif_step(x_seq[t], v, beta, theta, alpha)
        ss.append(s)
        us.append(u)
    return torch.stack(ss), v, torch.stack(us)


def pack_spike_mask(spikes):
    """Bit-pack a {0,1} tensor along its last dim (8 neurons / byte, LSB = lowest channel)."""
    s = spikes.reshape(-1, 8).to(torch.int32)
    w = (2 ** torch.arange(8, dtype=torch.int32, device=s.device))
    return (s * w).sum(-1).to(torch.uint8).reshape(*spikes.shape[:-1], spikes.shape[-1] // 8)


# ----------------------------------------------------------------------------------------------
# bf16 operand rounding (numeric contract of the tensor-core convs)
# ----------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x, on):
    return _RoundBF16.apply(x) if on else x


# ----------------------------------------------------------------------------------------------
# restated modules (same attribute / parameter names as the reference so state_dicts interchange)
# ----------------------------------------------------------------------------------------------
class OracleConvBlock(nn.Module):
    """reference model.py:9-18 (Conv2d(bias=False) -> BatchNorm2d -> SiLU) with SiLU|LIF."""

    def __init__(self, cin, cout, kernel_size=3, stride=1, padding=1, neuron="silu",
                 emulate_bf16=False, lif=None):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride, padding, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.neuron, self.emulate_bf16 = neuron, emulate_bf16
        self.lif = dict(LIF_DEFAULTS, **(lif or {}))

    def forward(self, x, v_prev=None):
        y = F.conv2d(_q(x, self.emulate_bf16), _q(self.conv.weight, self.emulate_bf16), None,
                     self.conv.stride, self.conv.padding)
        z = self.bn(y)
        if self.neuron == "silu":
            return F.silu(z), None
        if v_prev is None:
            v_prev = torch.zeros_like(z)
        s, v, u = lif_step(z, v_prev, self.lif["beta"], self.lif["v_th"], self.lif["alpha"])
        self.last_u = u.detach()      # pre-spike membrane, for the flip-rate protocol (|u - theta| < 1e-5)
        return s, v


class OracleDownBlock(nn.Module):
    """reference model.py:20-30."""

    def __init__(self, cin, cout, **kw):
        super().__init__()
        self.conv1 = OracleConvBlock(cin, cout, stride=2, **kw)
        self.conv2 = OracleConvBlock(cout, cout, **kw)

    def forward(self, x, v=None):
        v = v or (None, None)
        x, v1 = self.conv1(x, v[0])
        x, v2 = self.conv2(x, v[1])
        return x, (v1, v2)


class OracleUpBlock(nn.Module):
    """reference model.py:32-48."""

    def __init__(self, cin, cskip, cout, emulate_bf16=False, **kw):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
        self.conv1 = OracleConvBlock(cin // 2 + cskip, cout, emulate_bf16=emulate_bf16, **kw)
        self.conv2 = OracleConvBlock(cout, cout, emulate_bf16=emulate_bf16, **kw)
        self.emulate_bf16 = emulate_bf16

    def forward(self, x, skip, v=None):
        v = v or (None, None)
        x = F.conv_transpose2d(_q(x, self.emulate_bf16), _q(self.up.weight, self.emulate_bf16),
                               self.up.bias, stride=2)
        if x.shape[2:] != skip.shape[2:]:
            skip = F.interpolate(skip, size=x.shape[2:], mode="bilinear", align_corners=False)
        x = torch.cat([skip, x], dim=1)
        x, v1 = self.conv1(x, v[0])
        x, v2 = self.conv2(x, v[1])
        return x, (v1, v2)


class OracleConvLSTM2d(nn.Module):
    """reference model.py:50-71."""

    def __init__(self, cin, chid, kernel_size=3, emulate_bf16=False):
        super().__init__()
        self.hidden_channels = chid
        self.conv = nn.Conv2d(cin + chid, 4 * chid, kernel_size, padding=kernel_size // 2, bias=True)
        self.emulate_bf16 = emulate_bf16

    def forward(self, x, hidden_state=None):
        b, _, h, w = x.shape
        if hidden_state is None:
            hs = x.new_zeros(b, self.hidden_channels, h, w)
            cs = x.new_zeros(b, self.hidden_channels, h, w)
        else:
            hs, cs = hidden_state
        gates = F.conv2d(_q(torch.cat([x, hs], 1), self.emulate_bf16),
                         _q(self.conv.weight, self.emulate_bf16), self.conv.bias,
                         padding=self.conv.padding)
        i, f, g, o = torch.split(gates, self.hidden_channels, dim=1)
        c_next = torch.sigmoid(f) * cs + torch.sigmoid(i) * torch.tanh(g)
        h_next = torch.sigmoid(o) * torch.tanh(c_next)
        return h_next, (h_next, c_next)


class OracleTemporalUNet(nn.Module):
    """reference model.py:100-146 (ConvLSTM bottleneck variant), neuron = 'silu' | 'lif'.

    ``hidden_state`` is the opaque value threaded by the caller (train.py:62-66).  For 'silu' it is
    the reference's ``(h, c)``; for 'lif' it is ``((h, c), membranes)`` with one membrane tensor per
    ConvBlock.  ``widths`` lets tests build narrow copies (reference widths are the default).
    """

    def __init__(self, feature_channels, neuron="silu", emulate_bf16=False, lif=None,
                 widths=(128, 256, 512, 1024)):
        super().__init__()
        c3, c4, c5 = feature_channels
        w1, w2, w3, w4 = widths
        kw = dict(neuron=neuron, emulate_bf16=emulate_bf16, lif=lif)
        self.neuron, self.emulate_bf16 = neuron, emulate_bf16
        self.enc1, self.down1 = OracleConvBlock(c3, w1, **kw), OracleDownBlock(w1, w2, **kw)
        self.enc2, self.down2 = OracleConvBlock(w2 + c4, w2, **kw), OracleDownBlock(w2, w3, **kw)
        self.enc3, self.down3 = OracleConvBlock(w3 + c5, w3, **kw), OracleDownBlock(w3, w4, **kw)
        self.lstm = OracleConvLSTM2d(w4, w4, emulate_bf16=emulate_bf16)
        self.bottleneck_conv = OracleConvBlock(w4, w4, **kw)
        self.up1 = OracleUpBlock(w4, w3, w3, **kw)
        self.up2 = OracleUpBlock(w3, w2, w2, **kw)
        self.up3 = OracleUpBlock(w2, w1, w1, **kw)
        self.out_p5, self.out_p4, self.out_p3 = nn.Conv2d(w3, c5, 1), nn.Conv2d(w2, c4, 1), nn.Conv2d(w1, c3, 1)

    def _out(self, conv, x):
        return F.conv2d(_q(x, self.emulate_bf16), _q(conv.weight, self.emulate_bf16), conv.bias)

    def forward(self, features, hidden_state=None):
        p3, p4, p5 = features
        if self.neuron == "silu":
            lstm_state, m = hidden_state, {}
        else:
            lstm_state, m = hidden_state if hidden_state is not None else (None, {})
        nm = {}
        x1, nm["enc1"] = self.enc1(p3, m.get("enc1"))
        d, nm["down1"] = self.down1(x1, m.get("down1"))
        x2, nm["enc2"] = self.enc2(torch.cat([d, p4], 1), m.get("enc2"))
        d, nm["down2"] = self.down2(x2, m.get("down2"))
        x3, nm["enc3"] = self.enc3(torch.cat([d, p5], 1), m.get("enc3"))
        x, nm["down3"] = self.down3(x3, m.get("down3"))
        x, new_lstm = self.lstm(x, lstm_state)
        x, nm["bottleneck_conv"] = self.bottleneck_conv(x, m.get("bottleneck_conv"))
        d1, nm["up1"] = self.up1(x, x3, m.get("up1"))
        d2, nm["up2"] = self.up2(d1, x2, m.get("up2"))
        d3, nm["up3"] = self.up3(d2, x1, m.get("up3"))
        outs = (self._out(self.out_p3, d3), self._out(self.out_p4, d2), self._out(self.out_p5, d1))
        if self.neuron == "silu":
            return outs, new_lstm
        return outs, (new_lstm, nm)


def initialize_weights_oracle(m):
    """Restatement of reference weight_initialization.py:8-56 for the module types on the path."""
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, OracleConvLSTM2d):
        # applied AFTER the child Conv2d was visited (nn.Module.apply is children-first)
        nn.init.xavier_uniform_(m.conv.weight)
        nn.init.constant_(m.conv.bias, 0)
        n = m.conv.bias.size(0)
        m.conv.bias.data[n // 4:n // 2].fill_(1)


def run_sequence(net, feats_seq, hidden=None):
    """T-step unroll with state carry; returns last-step outputs (train.py:62-66) and state."""
    outs = None
    for feats in feats_seq:
        outs, hidden = net(feats, hidden)
    return outs, hidden


def train_step_oracle(net, feats_seq, optimizer, loss_fn, max_norm=10.0):
    """zero_grad -> unroll -> loss on last step -> backward -> clip 10 -> step (train.py:61-80)."""
    optimizer.zero_grad(set_to_none=True)
    outs, _ = run_sequence(net, feats_seq)
    loss = loss_fn(outs)
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=max_norm)
    optimizer.step()
    return loss.detach(), gn
