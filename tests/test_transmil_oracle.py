"""The numpy oracle of the TransMIL / Nystrom path against golden vectors produced by the reference itself
(tests/golden/make_golden_transmil.py), plus structural properties the GPU tests rely on."""
import numpy as np
import pytest

from conftest import golden_names, golden_x, load_golden
from oracle import transmil as O

RTOL, ATOL = 2e-4, 2e-5      # fp32 oracle vs fp32 reference: different summation orders only


@pytest.mark.parametrize("name", golden_names("nystrom_"))
def test_nystrom_attention_matches_reference(name):
    w, meta = load_golden(name)
    dim, dim_head, heads, m, ks, residual, iters = (int(v) for v in meta["meta_cfg"])
    x = golden_x(meta).numpy()
    for dtype, rtol, atol in ((np.float32, RTOL, ATOL), (np.float64, 5e-5, 5e-6)):
        y = O.nystrom_attention(w, x, heads=heads, dim_head=dim_head, num_landmarks=m, pinv_iterations=iters,
                                residual=bool(residual), dtype=dtype)
        assert y.shape == meta["out"].shape
        np.testing.assert_allclose(y, meta["out"], rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", golden_names("translayer_"))
def test_trans_layer_matches_reference(name):
    w, meta = load_golden(name)
    y = O.trans_layer(w, golden_x(meta).numpy(), prefix="")
    np.testing.assert_allclose(y, meta["out"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", golden_names("ppeg_"))
def test_ppeg_matches_reference(name):
    w, meta = load_golden(name)
    _, gh, gw = (int(v) for v in meta["meta_cfg"])
    y = O.ppeg(w, golden_x(meta).numpy(), gh, gw)
    np.testing.assert_allclose(y, meta["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("transmil_"))
def test_transmil_matches_reference(name):
    w, meta = load_golden(name)
    x = golden_x(meta).numpy()
    y = O.transmil_forward(w, x)
    np.testing.assert_allclose(y, meta["out"], rtol=1e-3, atol=1e-4)
    y64 = O.transmil_forward(w, x, dtype=np.float64)
    np.testing.assert_allclose(y64, meta["out"], rtol=1e-3, atol=1e-4)


def test_pinv_converges_to_the_inverse_of_a_softmax_matrix():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((1, 2, 16, 16)) * 0.3 + 4 * np.eye(16)
    a = np.exp(a) / np.exp(a).sum(-1, keepdims=True)
    z = O.moore_penrose_iter_pinv(a, iters=30)
    np.testing.assert_allclose(a @ z, np.broadcast_to(np.eye(16), a.shape), atol=1e-8)


def test_state_dict_and_init_parity_with_reference_fixture():
    """Same parameter names / shapes as the reference, and the same initial values under the same seed
    (tests/golden/make_golden_transmil.py used torch.manual_seed(46) before constructing the reference TransMIL)."""
    import torch
    from acmil_b200 import Struct
    from acmil_b200.transmil import TransMIL
    w, meta = load_golden("transmil_d64_n300")
    d_feat, d_inner, n_class = (int(v) for v in meta["meta_cfg"])
    torch.manual_seed(46)
    m = TransMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class))
    sd = m.state_dict()
    assert sorted(sd) == sorted(w)
    for k, v in sd.items():
        assert tuple(v.shape) == w[k].shape, k
        np.testing.assert_array_equal(v.numpy(), w[k], err_msg=k)


@pytest.mark.parametrize("name", golden_names("transmil_"))
def test_torch_port_matches_reference(name):
    """oracle/torch_port.transmil_forward is what bench.py times as the CPU baseline of this path."""
    import torch
    from oracle import torch_port as T
    w, meta = load_golden(name)
    p = {k: torch.from_numpy(v) for k, v in w.items()}
    with torch.no_grad():
        y = T.transmil_forward(p, golden_x(meta)).numpy()
    np.testing.assert_allclose(y, meta["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("seeded_transmil_"))
def test_torch_port_matches_seeded_reference_fixtures(name):
    """SURVEY KAT4 and BASELINE.json configs[2] at N = 50 000: reference logits with the weights given by seed (the
    fixture stores a digest of the reference's state_dict; acmil_b200's constructor must draw the same values)."""
    import hashlib
    import torch
    from acmil_b200 import Struct
    from acmil_b200.transmil import TransMIL
    from oracle import torch_port as T
    _, meta = load_golden(name)
    d_feat, d_inner, n_class = (int(v) for v in meta["meta_cfg"])
    torch.manual_seed(int(meta["meta_model_seed"]))
    sd = TransMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class)).state_dict()
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    assert h.hexdigest() == str(meta["meta_w_sha"])
    if int(meta["meta_x_shape"][1]) > 5000 and not bool(int(__import__("os").environ.get("ACMIL_SLOW_CPU_TESTS", "0"))):
        pytest.skip("N = 50k on the CPU port takes ~10 s per run: set ACMIL_SLOW_CPU_TESTS=1 (the GPU suite checks it)")
    with torch.no_grad():
        y = T.transmil_forward({k: v for k, v in sd.items()}, golden_x(meta)).numpy()
    np.testing.assert_allclose(y, meta["out"], rtol=1e-4, atol=1e-5)
