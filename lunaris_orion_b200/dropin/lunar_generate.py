"""Shim: put lunaris_orion_b200/dropin ahead of the reference on sys.path and the unmodified reference
train_hybrid.py (`from lunar_generate import LunarisCoreVAE`, train_hybrid.py:45) builds the B200-native VAE."""
from lunaris_orion_b200.lunar_generate import *  # noqa: F401,F403
from lunaris_orion_b200.lunar_generate import LunarisCoreVAE, Encoder, Decoder, ResBlock, SelfAttention2d, mish  # noqa: F401
