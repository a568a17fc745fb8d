"""architecture/nystrom_attention.py of the reference (== pip nystrom-attention 0.0.12): NystromAttention."""
from ..transmil import NystromAttention  # noqa: F401
