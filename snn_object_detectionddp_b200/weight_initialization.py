"""Weight initialisation with the reference's semantics (reference weight_initialization.py:8-83).

Same public names (`initialize_weights`, `initialize_model`) so main.py:147,151 can import them from here.
Conv / ConvTranspose: Kaiming-normal (fan_out, relu), zero bias; BatchNorm: (1, 0); ConvLSTM2d: Xavier-uniform
weights, zero bias with the forget-gate slice [n/4, n/2) set to 1.  Applied to `temporal_unet` only; the Detect
head keeps its constructor init and the feature extractor is frozen.

Runs on whatever device the parameters live on (plain torch RNG, one-off, not on the hot path).  When the
parameters are already attached to a flat ParamStore the writes go through the strided logical views, so the
physical [Cout][tap][Cin] buffers are updated in place.
"""
import torch.nn as nn

from .model import ConvLSTM2d


def initialize_weights(m):
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, ConvLSTM2d):
        # nn.Module.apply visits children first, so this overrides the Kaiming init of m.conv
        nn.init.xavier_uniform_(m.conv.weight)
        if m.conv.bias is not None:
            nn.init.constant_(m.conv.bias, 0)
            n = m.conv.bias.size(0)
            m.conv.bias.data[n // 4:n // 2].fill_(1)


def initialize_model(model):
    """Initialise `model.temporal_unet` (reference weight_initialization.py:62-83)."""
    model.temporal_unet.apply(initialize_weights)
    return model
