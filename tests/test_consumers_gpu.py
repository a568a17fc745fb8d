"""CLAM_SB / CLAM_MB / IBMIL on the GPU (through the C-ABI pool kernels) against the reference's own outputs.
Tolerance (north_star): logits and attention scores within 1e-3 relative."""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import golden_names, golden_x, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("clam_"))
def test_clam_golden(name):
    from acmil_b200 import Struct
    from acmil_b200.architecture.clam import CLAM_MB, CLAM_SB
    w, g = load_golden(name)
    d_feat, d_inner, n_class, gate, dropout, mb, label = (int(v) for v in g["meta_conf"])
    m = (CLAM_MB if mb else CLAM_SB)(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class), gate=bool(gate),
                                     dropout=bool(dropout))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    m = m.cuda().eval()
    x = golden_x(g).cuda()
    with torch.no_grad():
        logits = m(x)
        a_raw = m(x, attention_only=True)
        logits2, inst = m(x, torch.tensor([label], device="cuda"), instance_eval=True)
    np.testing.assert_allclose(a_raw.cpu().numpy(), g["eval_A_raw"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(logits.cpu().numpy(), g["eval_logits"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(logits2.cpu().numpy(), g["eval_logits"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(float(inst), float(g["eval_inst_loss"]), rtol=1e-3, atol=1e-5)
    # the small attention modules stand alone as well (clam.py:33,64): A [N, K]
    with torch.no_grad():
        h = torch.relu(torch.nn.functional.linear(x[0], m.attention_net[0].weight, m.attention_net[0].bias))
        A, h_back = m.attention_net[-1](h)
    assert h_back is h
    np.testing.assert_allclose(A.T.cpu().numpy(), g["eval_A_raw"], rtol=1e-3, atol=2e-5)


def test_clam_trains_without_dropout():
    """dropout=False: the training forward is the fused kernel and the backward reaches every bag-path parameter."""
    from acmil_b200 import Struct
    from acmil_b200.architecture.clam import CLAM_SB
    torch.manual_seed(3)
    m = CLAM_SB(Struct(D_feat=384, D_inner=128, n_class=2), dropout=False).cuda().train()
    x = torch.randn(1, 700, 384, device="cuda")
    logits, inst = m(x, torch.tensor([1], device="cuda"), instance_eval=True)
    (logits.sum() + inst).backward()
    for k, p in m.named_parameters():
        if k.startswith(("attention_net", "classifiers")) or k.startswith("instance_classifiers.1"):
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
    with pytest.raises(NotImplementedError):
        CLAM_SB(Struct(D_feat=384, D_inner=128, n_class=2), dropout=True).cuda().train()(x)


@pytest.mark.parametrize("name", golden_names("ibmil_"))
def test_ibmil_golden(name):
    from acmil_b200 import Struct
    from acmil_b200.architecture.ibmil import IBMIL
    w, g = load_golden(name)
    d_feat, d_inner, n_class, n_conf = (int(v) for v in g["meta_conf"])
    c_path = None
    if n_conf:
        tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False)
        np.save(tmp.name, w["confounder_feat"])
        c_path = [tmp.name]
    m = IBMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, c_path=c_path, c_learn=False),
              confounder_merge=str(g["meta_merge"]))
    if c_path:
        os.unlink(c_path[0])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    m = m.cuda().eval()
    with torch.no_grad():
        y, M, A = m(golden_x(g).cuda())
    np.testing.assert_allclose(y.cpu().numpy(), g["out_Y"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(M.cpu().numpy(), g["out_M"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(A.cpu().numpy(), g["out_A"], rtol=1e-3, atol=1e-7)
