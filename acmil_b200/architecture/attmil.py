"""architecture/attmil.py of the reference: AttentionGated, DAttention."""
from ..heads import AttentionGated, DAttention  # noqa: F401
