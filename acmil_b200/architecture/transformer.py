"""architecture/transformer.py of the reference: ACMIL_GA, ABMIL, Attention_Gated (raw [K, N] scores)."""
from ..heads import ABMIL, ACMIL_GA, Attention_Gated  # noqa: F401
