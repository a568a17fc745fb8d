"""acmil_b200 -- B200-native gated-attention MIL aggregation (drop-in for dazhangyu123/ACMIL's heads).

Compute lives in ``lib/libacmil_b200.so`` (hand-written sm_100a CUDA behind the C-ABI of
``include/acmil_b200.h``); this package is the host-side mirror of the reference's nn.Module interface.
"""
from . import _lib  # noqa: F401
from .gated_pool import GatedPool, GatedPoolResult, GatedPoolSpec  # noqa: F401
from .heads import (ABMIL, ACMIL_GA, Attention2, Attention_Gated, Attention_with_Classifier,  # noqa: F401
                    AttentionGated, Classifier_1fc, DAttention, DimReduction)
from .mha import ACMIL_MHA, MHA, MutiHeadAttention, MutiHeadAttention_modify  # noqa: F401
from .transmil import PPEG, NystromAttention, TransLayer, TransMIL  # noqa: F401
from .consumers import CLAM_MB, CLAM_SB, IBMIL, Attn_Net, Attn_Net_Gated  # noqa: F401
from .losses import diversity_loss  # noqa: F401
from .utils import Struct, set_seed  # noqa: F401

__version__ = "0.1.0"
