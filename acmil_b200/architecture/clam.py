"""architecture/clam.py of the reference: Attn_Net, Attn_Net_Gated, CLAM_SB, CLAM_MB."""
from ..consumers import CLAM_MB, CLAM_SB, Attn_Net, Attn_Net_Gated  # noqa: F401
