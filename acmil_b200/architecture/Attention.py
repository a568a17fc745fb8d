"""architecture/Attention.py of the reference (DTFD flavour: forward(x, isNorm=True))."""
from .. import heads as _h
from ..heads import Attention2, Attention_with_Classifier  # noqa: F401


class Attention_Gated(_h.Attention_Gated):
    def __init__(self, L=512, D=128, K=1):
        super().__init__(L, D, K, norm_default=True)
