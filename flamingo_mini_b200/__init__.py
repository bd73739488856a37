"""flamingo_mini_b200 — B200 (sm_100a) implementation of the flamingo-mini trainable hot path.

Drop-in replacements for the reference's ``PerceiverResampler`` and ``GatedCrossAttentionBlock`` /
``ModifiedLMBlock`` (same constructors, forward signatures and parameter names), backed by hand-written
tcgen05/TMA CUDA kernels in ``libflamingo_b200.so`` (C ABI in ``include/flamingo_b200.h``).
"""
from .gated_cross_attention import GatedCrossAttentionBlock, MaskedCrossAttention, ModifiedLMBlock
from .perceiver_resampler import PerceiverAttentionLayer, PerceiverResampler
from .utils import FeedForward, SquaredReLU

__all__ = ["GatedCrossAttentionBlock", "MaskedCrossAttention", "ModifiedLMBlock", "PerceiverAttentionLayer",
           "PerceiverResampler", "FeedForward", "SquaredReLU"]
__version__ = "0.1.0"
