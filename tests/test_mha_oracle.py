"""oracle/mha.py against the vectors the reference itself produced (tests/golden/make_golden_mha.py)."""
import numpy as np
import pytest

from conftest import golden_names, golden_x, load_golden
from oracle import mha as O


@pytest.mark.parametrize("name", golden_names("acmilmha_"))
def test_acmil_mha_oracle_matches_reference(name):
    w, g = load_golden(name)
    d_feat, d_inner, n_class, n_token, n_masked = (int(v) for v in g["meta_conf"])
    x = golden_x(g).numpy()
    ev = O.acmil_mha_forward(w, x, n_token)
    np.testing.assert_allclose(ev["attns"], g["eval_attns"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ev["sub"], g["eval_sub"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ev["slide"], g["eval_slide"], rtol=1e-4, atol=1e-5)
    rands = [g[f"rand_{i}"] for i in range(n_token)]
    tr = O.acmil_mha_forward(w, x, n_token, n_masked=n_masked, mask_drop=float(g["meta_mask_drop"]), rands=rands)
    assert np.array_equal(tr["attns"] == -1e9, g["train_attns"] == -1e9)      # the masked positions, bit-exact
    np.testing.assert_allclose(tr["attns"], g["train_attns"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(tr["sub"], g["train_sub"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(tr["slide"], g["train_slide"], rtol=1e-4, atol=1e-5)


def test_mha_oracle_matches_reference():
    w, g = load_golden("mha_n900")
    y = O.mha_forward(w, golden_x(g).numpy())
    np.testing.assert_allclose(y, g["out"], rtol=1e-4, atol=1e-5)


def test_module_parameter_names_and_init_match_reference():
    """Same state_dict keys / shapes as the reference, same initial values under the same seed (q aside: the fixture
    re-draws it, see make_golden_mha.py)."""
    import torch
    from acmil_b200 import ACMIL_MHA, Struct
    w, g = load_golden("acmilmha_k5_n1500")
    d_feat, d_inner, n_class, n_token, n_masked = (int(v) for v in g["meta_conf"])
    torch.manual_seed(61)
    m = ACMIL_MHA(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class), n_token=n_token, n_masked_patch=n_masked,
                  mask_drop=float(g["meta_mask_drop"]))
    sd = m.state_dict()
    assert sorted(sd) == sorted(w)
    for k, v in sd.items():
        assert tuple(v.shape) == w[k].shape, k
        if k != "q":
            np.testing.assert_array_equal(v.numpy(), w[k], err_msg=k)
