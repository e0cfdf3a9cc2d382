"""B200-native (sm_100a) hot path of the spiking temporal-U-Net detector.

Host side is Python/PyTorch (plumbing: memory, streams, autograd graph, torch.distributed); all compute
on the path runs in hand-written CUDA kernels behind the C ABI of ``libsnnb200.so``
(``include/snn_b200.h``).  There is no CPU or library fallback.
"""
__version__ = "0.1.0"
