"""Golden vectors for the ResNet18 patch encoder from the REFERENCE ITSELF: /root/reference/models.py imported in the build
container with `timm` stubbed (its top-level `import timm` is the only thing missing; the ResNet class needs torchvision's
BasicBlock only).  Weights are regenerated from seeds by the tests (tests/resnet_common.py); only outputs are stored.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_resnet.py
"""
import os
import sys
import types

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(OUT))
from resnet_common import CASES, make_images, randomize_bn  # noqa: E402

torch.set_num_threads(8)
timm = types.ModuleType("timm")
timm.models = types.ModuleType("timm.models")
timm.models.vision_transformer = types.ModuleType("timm.models.vision_transformer")
timm.models.vision_transformer.VisionTransformer = object
sys.modules.update({"timm": timm, "timm.models": timm.models, "timm.models.vision_transformer": timm.models.vision_transformer})
sys.path.insert(0, "/root/reference")
import models as ref_models  # noqa: E402

for name, wseed, bseed, iseed, b, size in CASES:
    torch.manual_seed(wseed)
    m = ref_models.resnet18(pretrained=False)          # models.py:80-88, classes=100 head
    randomize_bn(m, bseed).eval()
    x = make_images(iseed, b, size)
    with torch.no_grad():
        logits = m(x)
        m.class_classifier = torch.nn.Identity()       # what build_model does (models.py:201-204)
        feats = m(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), features=feats.numpy(), logits=logits.numpy(),
                        meta=np.array([wseed, bseed, iseed, b, size], dtype=np.int64),
                        conv1_w_sum=np.float64(m.conv1.weight.double().sum().item()))
    print(name, feats.shape, float(feats.abs().mean()), float(logits.abs().mean()))
