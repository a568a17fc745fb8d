"""ACMIL_MHA / MHA on the GPU (through the C-ABI: tcgen05 GEMM for the front layer, pool kernel for the attention) against
the reference's own outputs.  Tolerance (north_star): logits and attention scores within 1e-3 relative, masked positions
bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden_names, golden_x, load_golden

pytestmark = pytest.mark.gpu


def _model(w, g):
    from acmil_b200 import ACMIL_MHA, Struct
    d_feat, d_inner, n_class, n_token, n_masked = (int(v) for v in g["meta_conf"])
    m = ACMIL_MHA(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class), n_token=n_token, n_masked_patch=n_masked,
                  mask_drop=float(g["meta_mask_drop"]))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})      # reference checkpoint layout
    return m.cuda(), n_token


@pytest.mark.parametrize("name", golden_names("acmilmha_"))
def test_acmil_mha_eval_and_train_golden(name, monkeypatch):
    w, g = load_golden(name)
    m, n_token = _model(w, g)
    x = golden_x(g).cuda()
    m.eval()
    with torch.no_grad():
        sub, slide, attns = m(x)
    assert attns.shape == g["eval_attns"].shape and sub.shape == g["eval_sub"].shape
    np.testing.assert_allclose(attns.cpu().numpy(), g["eval_attns"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(sub.cpu().numpy(), g["eval_sub"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(slide.cpu().numpy(), g["eval_slide"], rtol=1e-3, atol=1e-5)
    # train mode: masking on, dropout off, the reference's own uniform draws (transformer.py:168) replayed in call order
    m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    draws = [torch.from_numpy(g[f"rand_{i}"]).cuda() for i in range(n_token)]
    calls = []

    def fake_rand(*shape, **kw):
        r = draws[len(calls)]
        assert tuple(shape) == tuple(r.shape), (shape, r.shape)      # same call shape as the reference
        calls.append(1)
        return r

    import acmil_b200.mha as M
    monkeypatch.setattr(M.torch, "rand", fake_rand)
    with torch.no_grad():
        sub, slide, attns = m(x)
    monkeypatch.undo()
    assert len(calls) == n_token
    a = attns.cpu().numpy()
    assert np.array_equal(a == -1e9, g["train_attns"] == -1e9)
    np.testing.assert_allclose(a, g["train_attns"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(sub.cpu().numpy(), g["train_sub"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(slide.cpu().numpy(), g["train_slide"], rtol=1e-3, atol=1e-5)


def test_mha_golden():
    from acmil_b200 import MHA, Struct
    w, g = load_golden("mha_n900")
    d_feat, d_inner, n_class = (int(v) for v in g["meta_conf"])
    m = MHA(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(golden_x(g).cuda())
    np.testing.assert_allclose(y.cpu().numpy(), g["out"], rtol=1e-3, atol=1e-5)


def test_acmil_mha_full_size_vs_oracle_and_gradients():
    """BASELINE config 2 size (N = 50k) against the numpy oracle, and the training surface: gradients of a train-mode
    forward against the same computation written with torch ops."""
    from acmil_b200 import ACMIL_MHA, Struct
    from oracle import mha as O
    torch.manual_seed(5)
    m = ACMIL_MHA(Struct(D_feat=384, D_inner=128, n_class=2), n_token=5, n_masked_patch=10, mask_drop=0.6)
    with torch.no_grad():
        m.q.normal_(0, 1.0)
    p = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    x = torch.randn(1, 50000, 384, generator=torch.Generator().manual_seed(6))
    ref = O.acmil_mha_forward(p, x.numpy(), 5)
    with torch.no_grad():
        sub, slide, attns = m(x.cuda())
    np.testing.assert_allclose(attns.cpu().numpy(), ref["attns"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(sub.cpu().numpy(), ref["sub"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(slide.cpu().numpy(), ref["slide"], rtol=1e-3, atol=1e-5)
    # gradients (eval mode: no mask, no dropout) versus a torch-op restatement of the reference forward
    xs = x[:, :3000].cuda()
    sub, slide, _ = m(xs)
    loss = (sub ** 2).sum() + slide.sum()
    loss.backward()
    got = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}
    m.zero_grad()
    import torch.nn.functional as F
    h = F.relu(F.linear(xs[0], m.dimreduction.fc1.weight))
    subs, pooled = [], []
    for i in range(5):
        att = m.sub_attention[i]
        u, c = att._score_vectors(m.q[:, i].unsqueeze(0))
        pr = torch.softmax(h @ u.T + c, dim=0).T @ h
        subs.append(m.classifier[i](att._finish(pr)))
        pooled.append(pr)
    loss2 = (torch.cat(subs) ** 2).sum() + m.Slide_classifier(m.bag_attention.from_pooled(torch.stack(pooled).mean(0))).sum()
    loss2.backward()
    for k, v in m.named_parameters():
        if v.grad is None:
            continue
        np.testing.assert_allclose(got[k].cpu().numpy(), v.grad.cpu().numpy(), rtol=2e-3, atol=1e-6, err_msg=k)
