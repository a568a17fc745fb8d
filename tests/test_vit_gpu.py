"""GPU parity of the ViT patch encoder (acmil_vit_fwd through the module) against the torchvision cross-check vectors and
the numpy oracle; tolerance = the fp16 feature store of Step2_feature_extract.py:165 (~5e-4 relative), tested at 1e-3."""
import numpy as np
import pytest
import torch

from conftest import golden_names, golden_x, load_golden
from vit_common import seeded_weights, timm_shapes

pytestmark = pytest.mark.gpu


def build(meta):
    from acmil_b200.vit import VisionTransformer
    img, patch, dim, depth, heads, mlp = (int(v) for v in meta["meta_cfg"])
    w = seeded_weights(timm_shapes(img, patch, dim, depth, mlp), int(meta["meta_w_seed"]))
    m = VisionTransformer(img_size=img, patch_size=patch, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=mlp / dim)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return m.cuda().eval(), w, heads, patch


@pytest.mark.parametrize("name", golden_names("vit_"))
def test_vit_matches_cross_check_vectors(name):
    _, meta = load_golden(name)
    m, _, _, _ = build(meta)
    with torch.no_grad():
        y = m(golden_x(meta).cuda()).cpu().numpy()
    assert y.shape == meta["out"].shape
    np.testing.assert_allclose(y, meta["out"], rtol=1e-3, atol=1e-4)


def test_custom_model_logits_and_features_vs_oracle():
    from acmil_b200 import Struct
    from acmil_b200.vit import CustomModel
    from oracle import vit as O
    _, meta = load_golden("vit_tiny_64")
    m, w, heads, patch = build(meta)
    torch.manual_seed(1)
    cm = CustomModel(Struct(n_class=4), m).cuda().eval()
    x = torch.randn(5, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    p = {"encoder." + k: v for k, v in w.items()}
    p["head.weight"], p["head.bias"] = cm.head.weight.detach().cpu().numpy(), cm.head.bias.detach().cpu().numpy()
    ref_logits, ref_feat = O.custom_model_forward(p, x.numpy(), num_heads=heads, patch=patch)
    with torch.no_grad():
        logits, feat = cm(x.cuda(), return_feature=True)
        only = cm(x.cuda())
    np.testing.assert_allclose(feat.cpu().numpy(), ref_feat, rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(logits.cpu().numpy(), ref_logits, rtol=1e-3, atol=1e-4)
    assert torch.equal(only, logits)


def test_vit_small_batch_independence_and_errors():
    from acmil_b200.vit import vit_small
    torch.manual_seed(3)
    m = vit_small(False, False, None).cuda().eval()
    x = torch.randn(9, 3, 224, 224, device="cuda")
    with torch.no_grad():
        y = m(x)
        y1 = m(x[4:5])
    assert y.shape == (9, 384) and bool(torch.isfinite(y).all())
    np.testing.assert_allclose(y[4:5].cpu().numpy(), y1.cpu().numpy(), rtol=1e-4, atol=1e-5)      # one image does not see the others
    with pytest.raises(AssertionError), torch.no_grad():
        m(torch.randn(1, 3, 192, 192, device="cuda"))
    with pytest.raises(NotImplementedError):
        m(x[:1])                                                    # grad mode with parameters that require grad
