"""architecture/ibmil.py of the reference: Attention_Gated, IBMIL."""
from ..consumers import IBMIL, Attention_Gated  # noqa: F401
