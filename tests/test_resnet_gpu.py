"""GPU parity of the ResNet18 patch encoder: im2col + tcgen05 GEMM path through the module mirror
against the reference's golden outputs and the numpy oracle.  Tolerance: 1e-3 relative (north_star) on the features."""
import os

import numpy as np
import pytest
import torch

from oracle import resnet as O
from resnet_common import CASES, make_images, np_state, randomize_bn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mirror(wseed, bseed):
    from acmil_b200.resnet import resnet18
    torch.manual_seed(wseed)
    return randomize_bn(resnet18(pretrained=False), bseed).eval()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_features_and_logits_match_reference(case):
    name, wseed, bseed, iseed, b, size = case
    g = np.load(os.path.join(GOLD, name + ".npz"))
    m = _mirror(wseed, bseed).cuda()
    x = make_images(iseed, b, size).cuda()
    from acmil_b200 import _lib
    before = _lib.launch_count()
    with torch.no_grad():
        logits = m(x)
        m.class_classifier = torch.nn.Identity()
        feats = m(x)
    assert _lib.launch_count() > before
    scale = np.abs(g["features"]).max()
    np.testing.assert_allclose(feats.cpu().numpy(), g["features"], rtol=1e-3, atol=1e-4 * scale)
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], rtol=1e-3, atol=1e-4 * np.abs(g["logits"]).max())


def test_against_oracle_and_chunking():
    """Batch larger than one chunk, odd image size (ragged im2col borders)."""
    m = _mirror(5, 6)
    m.CHUNK = 2
    x = make_images(7, 5, 72)
    ref, _ = O.resnet18_forward({k: v for k, v in np_state(m).items() if not k.startswith("class_classifier")}, x.numpy())
    m = m.cuda()
    m.class_classifier = torch.nn.Identity()
    with torch.no_grad():
        got = m(x.cuda()).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())


def test_build_model_dispatch_and_errors():
    from acmil_b200 import Struct
    from acmil_b200.resnet import ResNet
    import acmil_b200.resnet as R
    import acmil_b200.vit as V
    orig = R.resnet18
    try:
        R.resnet18 = lambda *a, **k: orig(pretrained=False)      # no network in the test box
        model = V.build_model(Struct(pretrain='natural_supervised', backbone='Resnet18', n_class=2)).cuda().eval()
    finally:
        R.resnet18 = orig
    assert isinstance(model.encoder, ResNet) and model.encoder.embed_dim == 512
    with torch.no_grad():
        logits, feats = model(make_images(1, 2, 64).cuda(), return_feature=True)
    assert logits.shape == (2, 2) and feats.shape == (2, 512)
    model.encoder.train()
    with pytest.raises(NotImplementedError):
        with torch.no_grad():
            model.encoder(make_images(1, 1, 64).cuda())
