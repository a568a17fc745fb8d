"""oracle/consumers.py against the vectors the reference itself produced (tests/golden/make_golden_consumers.py), and the
checkpoint layout of the drop-in modules."""
import numpy as np
import pytest

from conftest import golden_names, golden_x, load_golden
from oracle import consumers as O

CLAM_SEEDS = {"clam_sb_n1200": 71, "clam_sb_nogate_d512_n800": 73, "clam_mb_c3_n1000": 75}


@pytest.mark.parametrize("name", golden_names("clam_"))
def test_clam_oracle_matches_reference(name):
    w, g = load_golden(name)
    d_feat, d_inner, n_class, gate, dropout, mb, label = (int(v) for v in g["meta_conf"])
    r = O.clam_forward(w, golden_x(g).numpy(), bool(mb), bool(gate), att_index=3 if dropout else 2)
    np.testing.assert_allclose(r["A_raw"], g["eval_A_raw"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r["logits"], g["eval_logits"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", golden_names("ibmil_"))
def test_ibmil_oracle_matches_reference(name):
    w, g = load_golden(name)
    r = O.ibmil_forward(w, golden_x(g).numpy(), merge=str(g["meta_merge"]))
    np.testing.assert_allclose(r["Y"], g["out_Y"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r["M"], g["out_M"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r["A"], g["out_A"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("clam_"))
def test_clam_module_checkpoint_layout(name):
    """Same state_dict keys / shapes as the reference and, under the same seed, the same initial weights (biases aside:
    the fixture re-draws them, see make_golden_consumers.py)."""
    import torch
    from acmil_b200 import Struct
    from acmil_b200.architecture.clam import CLAM_MB, CLAM_SB
    w, g = load_golden(name)
    d_feat, d_inner, n_class, gate, dropout, mb, label = (int(v) for v in g["meta_conf"])
    torch.manual_seed(CLAM_SEEDS[name])
    m = (CLAM_MB if mb else CLAM_SB)(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class), gate=bool(gate),
                                     dropout=bool(dropout))
    sd = m.state_dict()
    assert sorted(sd) == sorted(w)
    for k, v in sd.items():
        assert tuple(v.shape) == w[k].shape, k
        if not k.endswith("bias"):
            np.testing.assert_array_equal(v.numpy(), w[k], err_msg=k)


def test_ibmil_module_checkpoint_layout():
    import torch
    from acmil_b200 import Struct
    from acmil_b200.architecture.ibmil import IBMIL
    w, g = load_golden("ibmil_n900")
    torch.manual_seed(77)
    m = IBMIL(Struct(D_feat=384, D_inner=128, n_class=2, c_path=None))
    sd = m.state_dict()
    assert sorted(sd) == sorted(w)
    for k, v in sd.items():
        np.testing.assert_array_equal(v.numpy(), w[k], err_msg=k)
