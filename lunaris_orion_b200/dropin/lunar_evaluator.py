"""Shim: `from lunar_evaluator import LunarMoETeacher` (reference train_hybrid.py:46) -> the B200-native Teacher."""
from lunaris_orion_b200.lunar_evaluator import *  # noqa: F401,F403
from lunaris_orion_b200.lunar_evaluator import (LunarMoETeacher, ExpertBlock, PixelArtAttention,  # noqa: F401
                                                 PixelArtFeatureExtractor, mish)
