"""The numpy ViT oracle against the torchvision cross-check vectors (tests/golden/make_golden_vit.py) and the host module's
parameter layout.  Parity with timm itself is unpinned (timm is not installable here) -- see oracle/vit.py."""
import numpy as np
import pytest

from conftest import golden_names, golden_x, load_golden
from oracle import vit as O
from vit_common import seeded_weights, timm_shapes


@pytest.mark.parametrize("name", golden_names("vit_"))
def test_vit_oracle_matches_torchvision(name):
    _, meta = load_golden(name)
    img, patch, dim, depth, heads, mlp = (int(v) for v in meta["meta_cfg"])
    w = seeded_weights(timm_shapes(img, patch, dim, depth, mlp), int(meta["meta_w_seed"]))
    y = O.vit_forward(w, golden_x(meta).numpy(), num_heads=heads, patch=patch)
    np.testing.assert_allclose(y, meta["out"], rtol=2e-4, atol=2e-5)


def test_vit_small_has_timm_parameter_layout():
    from acmil_b200.vit import CustomModel, vit_small
    from acmil_b200 import Struct
    m = vit_small(False, False, None)
    shapes = timm_shapes(224, 16, 384, 12, 1536)
    sd = m.state_dict()
    assert sorted(sd) == sorted(shapes)
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert sum(v.numel() for v in sd.values()) == 21_665_664      # ViT-S/16 without a head
    cm = CustomModel(Struct(n_class=3), m)
    assert tuple(cm.head.weight.shape) == (3, 384) and m.embed_dim == 384
