"""architecture/network.py of the reference: DimReduction, Classifier_1fc."""
from ..heads import Classifier_1fc, DimReduction  # noqa: F401
