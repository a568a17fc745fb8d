"""architecture/transformer.py of the reference: ACMIL_GA, ABMIL, Attention_Gated (raw [K, N] scores), ACMIL_MHA, MHA."""
from ..heads import ABMIL, ACMIL_GA, Attention_Gated  # noqa: F401
from ..mha import ACMIL_MHA, MHA, MutiHeadAttention, MutiHeadAttention_modify  # noqa: F401
