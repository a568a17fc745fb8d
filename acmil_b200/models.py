"""Import-swap module for the reference's ``models.py``: ``from models import build_model`` (Step2_feature_extract.py:13)
becomes ``from acmil_b200.models import build_model``.  Same names, constructors and forward signatures; the compute runs in
libacmil_b200.so (see vit.py / resnet.py)."""
from .resnet import ResNet, resnet18  # noqa: F401  (models.py:13-88)
from .vit import CustomModel, VisionTransformer, build_model, vit_small  # noqa: F401  (models.py:138-149, 166-215)
