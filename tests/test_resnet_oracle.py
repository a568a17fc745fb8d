"""CPU: the numpy ResNet18 oracle against the reference's own outputs (tests/golden/resnet18_*.npz, written by
tests/golden/make_golden_resnet.py from /root/reference/models.py), and the module mirror's construction parity
(same parameter names / shapes / seeded initial values as the reference's class)."""
import os

import numpy as np
import pytest
import torch

from oracle import resnet as O
from resnet_common import CASES, make_images, np_state, randomize_bn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mirror(wseed, bseed):
    from acmil_b200.resnet import resnet18
    torch.manual_seed(wseed)
    return randomize_bn(resnet18(pretrained=False), bseed).eval()


@pytest.mark.parametrize("case", [c for c in CASES if c[5] <= 96], ids=lambda c: c[0])
def test_oracle_matches_reference_golden(case):
    name, wseed, bseed, iseed, b, size = case
    g = np.load(os.path.join(GOLD, name + ".npz"))
    m = _mirror(wseed, bseed)
    # the mirror draws the same weights as the reference class did when the golden file was written
    assert abs(float(m.conv1.weight.detach().double().sum()) - float(g["conv1_w_sum"])) < 1e-9
    x = make_images(iseed, b, size).numpy()
    feats, logits = O.resnet18_forward(np_state(m), x)
    np.testing.assert_allclose(feats, g["features"], rtol=1e-3, atol=1e-4)      # north_star: 1e-3 fp32 rel-tol
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-3, atol=1e-4)
    f64, _ = O.resnet18_forward(np_state(m), x, dtype=np.float64)
    np.testing.assert_allclose(feats, f64, rtol=1e-3, atol=1e-4)


def test_mirror_state_dict_layout():
    m = _mirror(1, 2)
    keys = set(m.state_dict().keys())
    for k in ("conv1.weight", "bn1.running_var", "layer1.0.conv1.weight", "layer2.0.downsample.0.weight",
              "layer4.1.bn2.bias", "class_classifier.weight"):
        assert k in keys
    assert m.state_dict()["layer3.0.downsample.0.weight"].shape == (256, 128, 1, 1)
    assert sum(p.numel() for p in m.parameters()) == 11227812      # ResNet18 trunk + Linear(512, 100)


def test_cpu_input_raises():
    m = _mirror(1, 2)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m(torch.zeros(1, 3, 64, 64))


def test_fold_matches_conv_then_batchnorm():
    """acmil_b200.resnet._fold (eval BatchNorm folded into im2col-ordered weights + bias) against conv -> BN in torch."""
    import torch.nn.functional as F
    from acmil_b200.resnet import _fold
    torch.manual_seed(0)
    for cin, cout, k, stride, pad in [(3, 8, 7, 2, 3), (8, 16, 3, 1, 1), (8, 16, 1, 2, 0)]:
        conv = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)
        bn = randomize_bn(torch.nn.BatchNorm2d(cout), 5).eval()
        x = torch.randn(2, cin, 13, 11)
        with torch.no_grad():
            ref = bn(conv(x))
            w, b = _fold(conv, bn)
            # the column order of acmil_im2col: (ky, kx, c); unfold gives (c, ky, kx)
            cols = F.unfold(x, k, padding=pad, stride=stride)                       # [B, c*k*k, L]
            B, _, L = cols.shape
            cols = cols.view(B, cin, k, k, L).permute(0, 4, 2, 3, 1).reshape(B * L, k * k * cin)
            cols = F.pad(cols, (0, w.shape[1] - cols.shape[1]))
            out = (cols @ w.T + b).view(B, L, cout).permute(0, 2, 1).reshape(ref.shape)
        assert w.shape[1] % 4 == 0
        np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)
