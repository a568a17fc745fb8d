"""architecture/transMIL.py of the reference: TransLayer, PPEG, TransMIL."""
from ..transmil import PPEG, TransLayer, TransMIL  # noqa: F401
