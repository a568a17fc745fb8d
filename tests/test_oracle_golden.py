"""Pins the numpy oracle (oracle/gated_pool.py) to vectors produced by the reference
itself (tests/golden/make_golden.py).  CPU only.

Tolerances: the oracle is fp32 numpy (OpenBLAS) vs the reference's fp32 torch (MKL):
same arithmetic, different summation order -> a few ulp.  Mask indices: exact.
"""
import numpy as np
import pytest

from conftest import golden_names, golden_x, load_golden, np_seeded_state
from oracle import gated_pool as O

RTOL, ATOL = 2e-5, 2e-6


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", golden_names("acmil_ga_"))
def test_acmil_ga_eval(name):
    w, g = load_golden(name)
    x = golden_x(g).numpy()
    r = O.acmil_ga_forward(w, x)
    close(r["A_out"], g["eval_A"])
    close(r["sub"], g["eval_sub"])
    close(r["slide"], g["eval_slide"])
    close(r["bag_feat"], g["eval_feat"])
    close(O.acmil_ga_forward_feature(w, x), g["eval_feat"])
    close(O.branch_diversity_loss(r["A_out"]), g["eval_div"], rtol=1e-4)
    close(O.attention_entropy_loss(r["A_out"]), g["eval_ent"], rtol=1e-4)
    # bag_feat == afeat.mean(0) (probed property)
    close(r["bag_feat"][0], r["afeat"].mean(0), rtol=1e-5)
    # fp64 oracle agrees with the fp32 reference to fp32 noise
    r64 = O.acmil_ga_forward(w, x, dtype=np.float64)
    close(r64["A_out"], g["eval_A"], rtol=1e-4, atol=5e-6)


@pytest.mark.parametrize("name", [n for n in golden_names("acmil_ga_") if n != "acmil_ga_k1_n1024"])
def test_acmil_ga_train_mask(name):
    w, g = load_golden(name)
    x = golden_x(g).numpy()
    d_feat, d_inner, n_class, k, n_masked = (int(v) for v in g["meta_conf"])
    drop = float(g["meta_mask_drop"])
    r = O.acmil_ga_forward(w, x, training=True, n_masked_patch=n_masked, mask_drop=drop, rand=g["train_rand"])
    got = np.sort(r["masked_indices"], axis=-1)
    assert got.shape == g["train_masked_sorted"].shape
    assert np.array_equal(got, g["train_masked_sorted"]), "mask indices must be bit-exact"
    a_ref = g["train_A"]
    assert np.array_equal(r["A_out"] == O.MASK_FILL, a_ref == -1e9)
    close(r["A_out"], a_ref)
    close(r["sub"], g["train_sub"])
    close(r["slide"], g["train_slide"])
    close(r["bag_feat"], g["train_feat"])
    close(O.branch_diversity_loss(r["A_out"]), g["train_div"], rtol=1e-4)
    # masked positions carry exactly zero weight
    pr = O.softmax_rows(r["A_out"][0])
    assert np.all(pr[r["A_out"][0] == O.MASK_FILL] == 0)


def test_survey_kats():
    """SURVEY.md section 4 KAT1-3 literal numbers (independent of the .npz contents)."""
    w, g = load_golden("acmil_ga_k1_n1024")
    x = golden_x(g).numpy()
    r = O.acmil_ga_forward(w, x)
    close(r["slide"][0], [0.00249968, 0.11317079], rtol=1e-4, atol=1e-6)
    close(r["sub"][0], [-0.17061177, -0.04955589], rtol=1e-4, atol=1e-6)
    close(r["A_out"][0, 0, :4], [0.08671319, -0.06579139, 0.03284095, 0.11549069], rtol=1e-4, atol=1e-6)
    assert int(r["A_out"][0, 0].argmax()) == 697
    w, g = load_golden("acmil_ga_k5_n1024")
    r = O.acmil_ga_forward(w, x)
    close(r["slide"][0], [-0.07252882, 0.08611639], rtol=1e-4, atol=1e-6)
    assert O.topk_indices(r["A_out"][0], 10)[0].tolist() == [697, 423, 442, 824, 907, 912, 3, 685, 137, 870]
    r = O.acmil_ga_forward(w, x, training=True, n_masked_patch=10, mask_drop=0.6, rand=g["train_rand"])
    assert np.sort(r["masked_indices"], -1).tolist() == [
        [3, 137, 423, 697, 907, 912], [56, 85, 91, 98, 223, 717], [64, 139, 412, 468, 473, 965],
        [129, 449, 614, 822, 947, 953], [96, 189, 290, 606, 900, 914]]
    close(r["slide"][0], [-0.07270755, 0.08633027], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("abmil_"))
def test_abmil(name):
    w, g = load_golden(name)
    x = golden_x(g).numpy()
    close(O.abmil_forward(w, x)["out"], g["eval_out"])


@pytest.mark.parametrize("name", golden_names("attention_py_"))
def test_attention_py(name):
    w, g = load_golden(name)
    x = golden_x(g).numpy()

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in w.items() if k.startswith(prefix)}

    gw = sub("gate::")
    raw = O.attention_gated(x, gw["attention_V.0.weight"], gw["attention_V.0.bias"], gw["attention_U.0.weight"],
                            gw["attention_U.0.bias"], gw["attention_weights.weight"], gw["attention_weights.bias"])
    close(raw, g["gate_raw"])
    close(O.softmax_rows(raw), g["gate_norm"], rtol=1e-4, atol=1e-9)
    close(O.attention_with_classifier(sub("awc::"), x), g["awc_pred"])
    tw = sub("tgate::")
    close(O.attention_gated(x, tw["attention_V.0.weight"], tw["attention_V.0.bias"], tw["attention_U.0.weight"],
                            tw["attention_U.0.bias"], tw["attention_weights.weight"], tw["attention_weights.bias"]),
          g["tgate_raw"])


ATTMIL_AG_SHAPES = lambda bias: {  # noqa: E731  (state_dict order of attmil.AttentionGated)
    "feature.0.weight": (512, 1024), "feature.0.bias": (512,),
    "classifier.0.weight": (2, 512), "classifier.0.bias": (2,),
    "attention_a.0.weight": (128, 512), **({"attention_a.0.bias": (128,)} if bias else {}),
    "attention_b.0.weight": (128, 512), **({"attention_b.0.bias": (128,)} if bias else {}),
    "attention_c.weight": (1, 128), **({"attention_c.bias": (1,)} if bias else {}),
}
ATTMIL_DA_SHAPES = {  # state_dict order of attmil.DAttention(n_classes=3)
    "feature.0.weight": (512, 1024), "feature.0.bias": (512,),
    "attention.0.weight": (128, 512), "attention.0.bias": (128,),
    "attention.2.weight": (1, 128), "attention.2.bias": (1,),
    "classifier.0.weight": (3, 512), "classifier.0.bias": (3,),
}


def test_attmil():
    _, g = load_golden("attmil_n600")
    x = golden_x(g).numpy()
    seed = int(g["meta_w_seed"])
    for act in ("relu", "gelu", "tanh"):
        for bias in (False, True):
            p = np_seeded_state(ATTMIL_AG_SHAPES(bias), seed)
            close(O.attmil_attention_gated(p, x, act=act)["out"], g[f"ag_{act}_{int(bias)}"], rtol=1e-4, atol=1e-5)
    for act in ("relu", "gelu"):
        p = np_seeded_state(ATTMIL_DA_SHAPES, seed + 1)
        r = O.attmil_dattention(p, x, act=act)
        close(r["out"], g[f"da_{act}_y"], rtol=1e-4, atol=1e-5)
        close(r["A"], g[f"da_{act}_A"], rtol=1e-4, atol=1e-9)
        close(r["A_ori"], g[f"da_{act}_Aori"], rtol=1e-4, atol=1e-5)


def test_partials_merge_equals_softmax_pool():
    """Row-partition invariance that the sharded head relies on."""
    rng = np.random.default_rng(0)
    h = rng.standard_normal((1000, 64))
    a = rng.standard_normal((5, 1000)) * 3
    a[2, 17] = O.MASK_FILL
    ref = O.softmax_rows(a) @ h
    cuts = [0, 1, 130, 131, 700, 1000]
    parts = [O.pool_partials(h[s:e], a[:, s:e]) for s, e in zip(cuts[:-1], cuts[1:])]
    got, _, _ = O.merge_partials(*zip(*parts))
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name", golden_names("acmil_ga_"))
def test_torch_port_matches_golden(name):
    """oracle/torch_port.py (the CPU baseline bench.py times) against the reference's vectors."""
    import torch
    from oracle import torch_port as T
    w, g = load_golden(name)
    p = {k: torch.from_numpy(v) for k, v in w.items()}
    x = golden_x(g)
    with torch.no_grad():
        sub, slide, a = T.acmil_ga_forward(p, x)
    close(a.numpy(), g["eval_A"])
    close(sub.numpy(), g["eval_sub"])
    close(slide.numpy(), g["eval_slide"])
    if "train_rand" in g and g["train_masked_sorted"].size:
        n_masked = int(g["meta_conf"][4])
        with torch.no_grad():
            sub, slide, a = T.acmil_ga_forward(p, x, True, n_masked, float(g["meta_mask_drop"]),
                                               torch.from_numpy(g["train_rand"]))
        assert np.array_equal(a.numpy() == -1e9, g["train_A"] == -1e9)
        close(slide.numpy(), g["train_slide"])
