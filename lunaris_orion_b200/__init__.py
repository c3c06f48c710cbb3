"""lunaris_orion_b200 — B200-native (sm_100a) hot path of Lunaris-Orion's hybrid VAE + Teacher training step.

Host side is Python/PyTorch (device memory, streams, autograd graph, torch.distributed); all arithmetic on the
hot path is hand-written CUDA behind the C ABI in include/lunaris_b200.h (liblunaris_b200.so).
There is no CPU or library fallback: importing the ops without the built library raises.
"""
__version__ = "0.1.0"
