"""Parity of the CUDA path (through the C-ABI) against (1) golden vectors produced by the reference
itself and (2) the numpy oracle on seeded inputs up to the BASELINE.json sizes.

Tolerances (north_star: "slide logits and attention scores within 1e-3 fp32 rel-tol; top-k mask
indices bit-exact under fixed seed"):
  logits / pooled features : rtol 1e-3 (+ atol 1e-5)
  raw attention scores      : rtol 1e-3 + atol 1e-5   (|A| ~ 0.1; atol guards the zero crossings)
  mask / top-k indices      : exact
The kernels are fp32-faithful (error-compensated), so the tests also assert a 20x tighter bound to
catch regressions early; the official bound is the one above.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, golden_x, load_golden, np_seeded_state
from oracle import gated_pool as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-5
TIGHT = dict(rtol=5e-5, atol=2e-6)


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def both(a, b):
    close(a, b)
    close(a, b, **TIGHT)


def impls():
    import acmil_b200._lib as L
    return [L.IMPL_FFMA, L.IMPL_AUTO]


def make_acmil(w, g, impl):
    from acmil_b200 import ACMIL_GA, Struct
    d_feat, d_inner, n_class, k, n_masked = (int(v) for v in g["meta_conf"])
    conf = Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, n_token=k)
    m = ACMIL_GA(conf, n_token=k, n_masked_patch=n_masked, mask_drop=float(g["meta_mask_drop"]))
    m.load_state_dict({k_: torch.from_numpy(v) for k_, v in w.items()})   # reference checkpoint layout
    m._op.impl = impl
    return m.cuda()


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("name", golden_names("acmil_ga_"))
def test_acmil_ga_eval_golden(name, impl):
    w, g = load_golden(name)
    x = golden_x(g).cuda()
    m = make_acmil(w, g, impl).eval()
    with torch.no_grad():
        sub, slide, a = m(x)
        feat = m.forward_feature(x)
    assert sub.shape == g["eval_sub"].shape and slide.shape == g["eval_slide"].shape and a.shape == g["eval_A"].shape
    both(a, g["eval_A"])
    both(sub, g["eval_sub"])
    both(slide, g["eval_slide"])
    both(feat, g["eval_feat"])


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("name", [n for n in golden_names("acmil_ga_") if n != "acmil_ga_k1_n1024"])
def test_acmil_ga_train_mask_golden(name, impl):
    """Mask indices must be bit-exact given the reference's own uniform draw (stored in the fixture)."""
    w, g = load_golden(name)
    x = golden_x(g).cuda()
    m = make_acmil(w, g, impl).train()
    k, n_masked = int(g["meta_conf"][3]), int(g["meta_conf"][4])
    n = x.shape[1]
    nm = min(n_masked, n)
    keep = int(nm * float(g["meta_mask_drop"]))
    rsel = torch.argsort(torch.from_numpy(g["train_rand"]), dim=-1)[:, :keep].cuda()
    with torch.no_grad():
        branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
        head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
        res, _ = m._pool(x[0], n_masked=n_masked if keep else 0, keep=keep, rsel=rsel if keep else None,
                         branch=branch, head=head, slide_head=True)
    a = res.scores.cpu().numpy()
    if keep:
        got = np.sort(res.masked_idx[0].cpu().numpy(), axis=-1)
        assert np.array_equal(got, g["train_masked_sorted"]), "masked indices differ from the reference"
        # torch.topk order of the raw scores
        raw = g["eval_A"][0]
        assert np.array_equal(res.topk_idx[0, :, :nm].cpu().numpy(), O.topk_indices(raw, nm))
    assert np.array_equal(a == -1e9, g["train_A"][0] == -1e9)
    both(a, g["train_A"][0])
    both(res.sub[0], g["train_sub"])
    both(res.slide, g["train_slide"])
    both(res.bag_feat, g["train_feat"])


def test_module_rng_stream_matches_reference_call_sequence():
    """ACMIL_GA.forward in train mode must consume the generator exactly like transformer.py:316
    (one torch.rand(K, nm, device=x.device)), so a seeded run masks what the reference would mask."""
    w, g = load_golden("acmil_ga_k5_n1024")
    x = golden_x(g).cuda()
    m = make_acmil(w, g, 0).train()
    with torch.no_grad():
        m.eval()
        _, _, raw = m(x)
        m.train()
        torch.manual_seed(123)
        expect_rand = torch.rand(5, 10, device="cuda")
        after_ref = torch.rand(3, device="cuda")
        torch.manual_seed(123)
        _, _, a = m(x)
        after_mine = torch.rand(3, device="cuda")
    assert torch.equal(after_ref, after_mine), "generator stream diverged"
    top = torch.topk(raw[0], 10, dim=-1).indices
    rsel = torch.argsort(expect_rand, dim=-1)[:, :6]
    expect = torch.gather(top, 1, rsel)
    got = (a[0] == -1e9).nonzero()[:, 1].reshape(5, 6)
    assert torch.equal(got.sort(dim=-1).values, expect.sort(dim=-1).values)
    # eval never masks
    m.eval()
    with torch.no_grad():
        _, _, a2 = m(x)
    assert not bool((a2 == -1e9).any())


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("name", golden_names("abmil_"))
def test_abmil_golden(name, impl):
    from acmil_b200 import ABMIL, Struct
    w, g = load_golden(name)
    d_feat, d_inner, n_class = (int(v) for v in g["meta_conf"][:3])
    m = ABMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, n_token=1))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    m._op.impl = impl
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(golden_x(g).cuda())
    assert y.shape == g["eval_out"].shape
    both(y, g["eval_out"])


@pytest.mark.parametrize("name", golden_names("attention_py_"))
def test_attention_py_golden(name):
    from acmil_b200.architecture import Attention as A
    from acmil_b200.architecture.transformer import Attention_Gated as TG
    w, g = load_golden(name)
    Lw, D, K, ncls = (int(v) for v in g["meta_conf"])
    x = golden_x(g).cuda()

    def sub(prefix):
        return {k[len(prefix):]: torch.from_numpy(v) for k, v in w.items() if k.startswith(prefix)}

    gate = A.Attention_Gated(Lw, D, K)
    gate.load_state_dict(sub("gate::"))
    gate = gate.cuda().eval()
    awc = A.Attention_with_Classifier(Lw, D, K, ncls)
    awc.load_state_dict(sub("awc::"))
    awc = awc.cuda().eval()
    tg = TG(Lw, D, K)
    tg.load_state_dict(sub("tgate::"))
    tg = tg.cuda().eval()
    with torch.no_grad():
        both(gate(x, isNorm=False), g["gate_raw"])
        close(gate(x), g["gate_norm"], rtol=1e-3, atol=1e-9)
        both(awc(x), g["awc_pred"])
        both(tg(x), g["tgate_raw"])


def test_attmil_golden():
    from acmil_b200.architecture.attmil import AttentionGated, DAttention
    from test_oracle_golden import ATTMIL_AG_SHAPES, ATTMIL_DA_SHAPES
    _, g = load_golden("attmil_n600")
    x = golden_x(g).cuda()
    seed = int(g["meta_w_seed"])
    for act in ("relu", "gelu", "tanh"):
        for bias in (False, True):
            m = AttentionGated(act=act, bias=bias)
            m.load_state_dict({k: torch.from_numpy(v) for k, v in np_seeded_state(ATTMIL_AG_SHAPES(bias), seed).items()})
            m = m.cuda().eval()
            with torch.no_grad():
                close(m(x), g[f"ag_{act}_{int(bias)}"], rtol=1e-3, atol=2e-5)
    for act in ("relu", "gelu"):
        m = DAttention(n_classes=3, dropout=False, act=act)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in np_seeded_state(ATTMIL_DA_SHAPES, seed + 1).items()})
        m = m.cuda().eval()
        with torch.no_grad():
            y, a = m(x, return_attn=True)
            _, a_ori = m(x, return_attn=True, no_norm=True)
        close(y, g[f"da_{act}_y"], rtol=1e-3, atol=2e-5)
        close(a, g[f"da_{act}_A"], rtol=1e-3, atol=1e-9)
        close(a_ori, g[f"da_{act}_Aori"], rtol=1e-3, atol=2e-5)


# ----------------------------------------------------------------------------- oracle at full size
def _random_acmil(seed, d_feat=384, d_inner=128, k=5, n_class=2, n_masked=10, drop=0.6, scale_ww=1.0):
    from acmil_b200 import ACMIL_GA, Struct
    torch.manual_seed(seed)
    m = ACMIL_GA(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, n_token=k), n_token=k,
                 n_masked_patch=n_masked, mask_drop=drop)
    with torch.no_grad():
        m.attention.attention_weights.weight.mul_(scale_ww)
    return m


def _np_state(m):
    return {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("n,scale_ww", [(50000, 1.0), (50000, 40.0), (12345, 8.0)])
def test_full_size_vs_oracle(n, scale_ww, impl):
    """BASELINE.json config 2 size (N=50k, D=384, K=5, mask 10/0.6).  scale_ww > 1 makes the attention
    peaky (trained-model-like): the masked rows then carry most of the softmax mass, which is the case a
    subtract-after-the-fact implementation would get wrong."""
    m = _random_acmil(100 + n % 7, scale_ww=scale_ww)
    m._op.impl = impl
    p = _np_state(m)
    m = m.cuda()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(1, n, 384, generator=g)
    rand = torch.rand(5, 10, generator=g)
    ref32 = O.acmil_ga_forward(p, x.numpy(), training=True, n_masked_patch=10, mask_drop=0.6, rand=rand.numpy())
    ref64 = O.acmil_ga_forward(p, x.numpy(), training=True, n_masked_patch=10, mask_drop=0.6, rand=rand.numpy(),
                               dtype=np.float64)
    rsel = torch.argsort(rand, dim=-1)[:, :6].cuda()
    with torch.no_grad():
        branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
        head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
        res, _ = m._pool(x[0].cuda(), n_masked=10, keep=6, rsel=rsel, branch=branch, head=head, slide_head=True)
    got_mask = np.sort(res.masked_idx[0].cpu().numpy(), -1)
    ref_mask = np.sort(ref32["masked_indices"], -1)
    if not np.array_equal(got_mask, ref_mask):
        # only acceptable if the fp32 reference itself disagrees with fp64 there (a rounding-noise tie)
        assert not np.array_equal(ref_mask, np.sort(ref64["masked_indices"], -1)), "mask differs from a stable reference"
        pytest.skip("fp32 tie in the reference ordering for this seed")
    a = res.scores.cpu().numpy()
    close(a, ref32["A_out"][0])
    close(res.slide, ref32["slide"])
    close(res.sub[0], ref32["sub"])
    close(res.bag_feat, ref32["bag_feat"])
    # versus fp64 truth we must stay within a small multiple of the fp32 reference's own error (the tcgen05
    # kernel's split-fp16 products carry ~2^-21 and its gate uses ex2/rcp.approx: a few ulp more than fp32 libm)
    err_mine = np.abs(a - ref64["A_out"][0]).max()
    err_ref = np.abs(ref32["A_out"][0] - ref64["A_out"][0]).max()
    assert err_mine <= max(8 * err_ref, 2e-6), (err_mine, err_ref)


@pytest.mark.parametrize("impl", impls())
def test_ragged_batch_equals_single_bags(impl):
    """Several bags of different sizes in one launch == each bag alone (and empty-tail tiles are inert)."""
    m = _random_acmil(7).cuda().eval()
    m._op.impl = impl
    sizes = [1, 63, 64, 65, 127, 128, 129, 1000, 4097]
    g = torch.Generator().manual_seed(5)
    bags = [torch.randn(n, 384, generator=g).cuda() for n in sizes]
    off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    w = m._weights()
    op = m._op
    packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
    head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
    with torch.no_grad():
        res = op.run(packed, torch.cat(bags), off, branch_w=branch[0], branch_b=branch[1], head_w=head[0],
                     head_b=head[1], slide_head=True)
        for i, b in enumerate(bags):
            sub, slide, a = m(b[None])
            close(res.sub[i], sub.cpu().numpy(), rtol=1e-5, atol=1e-6)
            close(res.slide[i:i + 1], slide.cpu().numpy(), rtol=1e-5, atol=1e-6)
            close(res.scores[:, off[i]:off[i + 1]], a[0].cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_row_sharded_partials_merge(ranks, impl):
    """Bag sharded by rows over `ranks` shards (here: sequentially on one GPU), records concatenated the way
    all_gather_into_tensor would, finished once: must equal the unsharded result, masks included."""
    m = _random_acmil(11, scale_ww=20.0).cuda().train()
    m._op.impl = impl
    n = 10007
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, 384, generator=g).cuda()
    rsel = torch.argsort(torch.rand(5, 10, generator=g), dim=-1)[:, :6].cuda()
    w = m._weights()
    op = m._op
    packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
    head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
    kw = dict(keep=[6], rsel=rsel, branch_w=branch[0], branch_b=branch[1], head_w=head[0], head_b=head[1], slide_head=True)
    with torch.no_grad():
        full = op.run(packed, x, [0, n], n_masked=10, **kw)
        bounds = [n * r // ranks for r in range(ranks + 1)]
        bounds[1] = min(bounds[1], 3) if ranks == 8 else bounds[1]     # a shard smaller than n_masked
        recs, ctxs = [], []
        for r in range(ranks):
            xs = x[bounds[r]:bounds[r + 1]].contiguous()
            rec, ctx = op.partial(packed, xs, [0, xs.shape[0]], n_masked=10, shard_begin=[bounds[r]])
            recs.append(rec)
            ctxs.append(ctx)
        gathered = torch.cat(recs)
        outs = [op.finish(ctxs[r], gathered, ranks, **kw) for r in range(ranks)]
    for r, o in enumerate(outs):
        assert torch.equal(o.masked_idx.sort(-1).values, full.masked_idx.sort(-1).values)
        assert torch.equal(o.topk_idx, full.topk_idx)
        close(o.sub, full.sub.cpu().numpy(), rtol=2e-5, atol=1e-6)
        close(o.slide, full.slide.cpu().numpy(), rtol=2e-5, atol=1e-6)
        close(o.scores, full.scores[:, bounds[r]:bounds[r + 1]].cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_attn_stats_and_softmax_rows():
    from acmil_b200 import GatedPool
    w, g = load_golden("acmil_ga_k5_n1024")
    x = golden_x(g).cuda()
    m = make_acmil(w, g, 0).eval()
    with torch.no_grad():
        res, _ = m._pool(x[0])
        gram, ent, div = GatedPool.attn_stats(res.scores, [0, 1024], res.lse_m, res.lse_l)
        close(div, [float(g["eval_div"])], rtol=1e-3, atol=1e-7)
        close(ent.sum() / 5, float(g["eval_ent"]), rtol=1e-3)
        a = torch.from_numpy(g["train_A"][0]).cuda()       # with -1e9 entries
        sm = GatedPool.softmax_rows(a)
        close(sm, torch.softmax(a, dim=1).cpu().numpy(), rtol=1e-5, atol=1e-12)
        lse = torch.logsumexp(a, dim=1)
        mx = a.max(dim=1).values
        _, _, div2 = GatedPool.attn_stats(a, [0, 1024], mx[None].contiguous(), torch.exp(lse - mx)[None].contiguous())
        close(div2, [float(g["train_div"])], rtol=1e-3, atol=1e-7)


def test_backward_matches_torch_autograd():
    """Training: gradients of (sub CE + slide CE + diversity) w.r.t. every parameter equal those of the
    same graph written with torch ops (the reference's op sequence) on the same device and mask."""
    import torch.nn.functional as F
    m = _random_acmil(21).cuda().train()
    n = 3000
    x = torch.randn(1, n, 384, generator=torch.Generator().manual_seed(4)).cuda()
    y = torch.tensor([1], device="cuda")
    torch.manual_seed(77)
    sub, slide, a = m(x)
    masked = (a[0] == -1e9)
    assert int(masked.sum()) == 30

    def losses(sub, slide, a):
        p = torch.softmax(a, dim=-1)
        d = sum(torch.cosine_similarity(p[:, i], p[:, j], dim=-1).mean() for i in range(5) for j in range(i + 1, 5)) / 10
        return F.cross_entropy(sub, y.repeat_interleave(5)) + F.cross_entropy(slide, y) + d

    losses(sub, slide, a).backward()
    mine = {k: v.grad.clone() for k, v in m.named_parameters()}
    m.zero_grad()
    h = F.relu(F.linear(x[0], m.dimreduction.fc1.weight))
    g = m.attention
    s = F.linear(torch.tanh(g.attention_V[0](h)) * torch.sigmoid(g.attention_U[0](h)), g.attention_weights.weight,
                 g.attention_weights.bias).t().masked_fill(masked, -1e9)
    af = torch.softmax(s, 1) @ h
    sub2 = torch.stack([c.fc(af[i]) for i, c in enumerate(m.classifier)])
    slide2 = m.Slide_classifier.fc(torch.softmax(s, 1).mean(0, keepdim=True) @ h)
    close(sub, sub2.detach().cpu().numpy(), rtol=1e-4, atol=1e-6)
    losses(sub2, slide2, s[None]).backward()
    for k, v in m.named_parameters():
        close(mine[k], v.grad.cpu().numpy(), rtol=2e-3, atol=1e-6)


def test_cabi_error_behaviour():
    import ctypes as C
    import acmil_b200._lib as L
    lib = L.load()
    shape = L.GpShape(384, 128, 64, 5, 1, 0, 1, 0, 1, 1, 1, 0)      # d_attn 64: unsupported
    n = C.c_size_t(0)
    assert lib.acmil_gp_packed_bytes(C.byref(shape), C.byref(n)) == -4
    assert b"d_attn" in lib.acmil_last_error()
    m = _random_acmil(3).cuda().eval()
    with pytest.raises(ValueError):
        m._op.run(torch.empty(16, dtype=torch.uint8, device="cuda"), torch.zeros(4, 100, device="cuda"), [0, 4])
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 384))      # CPU tensor: no CPU path
    before = L.launch_count()
    with torch.no_grad():
        m(torch.zeros(1, 4, 384, device="cuda"))
    assert L.launch_count() >= before + 3


@pytest.mark.parametrize("n", [50000, 7])
def test_in_kernel_argsort_of_the_mask_draw_matches_torch(n):
    """acmil_gp_finish_rand (raw uniform draws, argsort inside the kernel) == acmil_gp_finish fed torch.argsort(draws)."""
    m = _random_acmil(3).cuda()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, 384, generator=g).cuda()
    nm = min(10, n)
    keep = int(nm * 0.6)
    rand = torch.rand(5, nm, generator=g)
    rsel = torch.argsort(rand, dim=-1)[:, :keep].cuda()
    branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
    head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
    with torch.no_grad():
        a, _ = m._pool(x, n_masked=10, keep=keep, rsel=rsel, branch=branch, head=head, slide_head=True)
        b, _ = m._pool(x, n_masked=10, keep=keep, rand=rand[None].cuda(), branch=branch, head=head, slide_head=True)
    assert torch.equal(a.masked_idx, b.masked_idx)          # same rows in the same order
    assert torch.equal(a.topk_idx, b.topk_idx)
    np.testing.assert_allclose(b.slide.cpu().numpy(), a.slide.cpu().numpy(), rtol=1e-5, atol=1e-6)


def _run_train_bags(m, xs, rand, impl):
    """Train-mode forward of several bags in one launch through the C-ABI -> (result, rescued flags per bag)."""
    op = m._op
    op.impl = impl
    w = m._weights()
    packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    sizes = [x.shape[0] for x in xs]
    off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    xcat = torch.cat(xs).cuda()
    keep = [int(min(10, n) * 0.6) for n in sizes]
    part, ctx = op.partial(packed, xcat, off, n_masked=10)
    res = op.finish(ctx, part, 1, keep=keep, rand=rand.cuda(),
                    branch_w=torch.stack([c.fc.weight for c in m.classifier]),
                    branch_b=torch.stack([c.fc.bias for c in m.classifier]),
                    head_w=m.Slide_classifier.fc.weight, head_b=m.Slide_classifier.fc.bias, slide_head=True)
    return res, op.rescued_bags(ctx)


@pytest.mark.parametrize("order", ["ascending", "descending", "equal", "block", "random"])
def test_sorted_bag_does_not_nan(order):
    """Score orders that defeat a bounded candidate scratch: rows sorted by ascending branch-0 score (every row is a new
    top-n member when it arrives), a contiguous block of > 256 high-score rows, all-equal scores.  The tcgen05 kernel
    flags such a bag and the exact FFMA kernel redoes it on the device; masks stay bit-exact, nothing is NaN, and the
    other bags of the same launch are untouched (the flag is per bag)."""
    import acmil_b200._lib as L
    n = 50000
    m = _random_acmil(3)
    p = _np_state(m)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, 384, generator=g)
    s0 = O.acmil_ga_forward(p, x[None].numpy())["A_out"][0][0]      # eval scores of branch 0
    if order == "ascending":
        x = x[torch.from_numpy(np.argsort(s0, kind="stable"))]
    elif order == "descending":
        x = x[torch.from_numpy(np.argsort(-s0, kind="stable"))]
    elif order == "equal":
        x = x[:1].repeat(n, 1).contiguous()
    elif order == "block":       # the 700 best rows of branch 0, ascending, in the middle of the bag
        idx = np.argsort(s0, kind="stable")
        top, rest = idx[-700:], np.sort(idx[:-700])
        x = x[torch.from_numpy(np.concatenate([rest[:20000], top, rest[20000:]]))]
    other = torch.randn(3000, 384, generator=g)                     # a second, ordinary bag in the same launch
    rand = torch.rand(2, 5, 10, generator=g)
    res, rescued = _run_train_bags(m, [x, other], rand, L.IMPL_AUTO)
    for b, xb in enumerate([x, other]):
        ref = O.acmil_ga_forward(p, xb[None].numpy(), training=True, n_masked_patch=10, mask_drop=0.6, rand=rand[b].numpy())
        sl = slice(res.row_offsets[b], res.row_offsets[b + 1])
        assert torch.isfinite(res.slide[b]).all() and torch.isfinite(res.sub[b]).all()
        got = np.sort(res.masked_idx[b].cpu().numpy(), -1)
        assert np.array_equal(got, np.sort(ref["masked_indices"], -1)), (order, b)
        close(res.scores[:, sl], ref["A_out"][0])
        close(res.slide[b:b + 1], ref["slide"])
        close(res.sub[b], ref["sub"])
    assert rescued[1] == 0
    if order in ("ascending", "block"):
        assert rescued[0] == 1, "expected the bounded scratch to overflow on this order (is the test still adversarial?)"
    if order in ("descending", "random", "equal"):
        assert rescued[0] == 0


class _LocalExchange:
    """acmil_gp_exchange whose "peers" are buffers of the same GPU: the in-kernel exchange protocol (records stored into
    every rank's gather buffer, per-source flags, epoch parity) exercised without a second device.  Same interface as
    acmil_b200.sharding.PeerExchange."""

    def __init__(self, world, rank, bufs, capacity):
        self.world, self.rank, self.bufs, self.capacity = world, rank, bufs, capacity
        self.state = torch.zeros(8, dtype=torch.int32, device="cuda")

    def c_struct(self, partial_bytes):
        import acmil_b200._lib as L
        assert partial_bytes <= self.capacity
        x = L.GpExchange()
        x.n_ranks, x.rank = self.world, self.rank
        for r, b in enumerate(self.bufs):
            x.d_flags[r] = b.data_ptr()
            x.d_gather[r] = b.data_ptr() + 256
        x.d_epoch = self.state.data_ptr()
        x.d_ticket = self.state.data_ptr() + 16
        x.gather_bytes = 2 * self.world * self.capacity
        return x


@pytest.mark.parametrize("impl", impls())
@pytest.mark.parametrize("ranks", [1, 2, 8])
def test_in_kernel_exchange_matches_unsharded(ranks, impl):
    """The fused exchange (acmil_gp_partial_x / acmil_gp_finish_x): every "rank" pushes its records into all gather
    buffers and raises its flag, every rank's finish kernel waits for the flags and must reproduce the unsharded result,
    masks included -- over three consecutive steps (both parities of the double-buffered gather area, epoch advance)."""
    m = _random_acmil(13, scale_ww=10.0).cuda().train()
    m._op.impl = impl
    sizes = [6007, 3001]
    g = torch.Generator().manual_seed(21)
    w = m._weights()
    op = m._op
    packed = op.pack(w.get("w1"), None, w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    branch = (torch.stack([c.fc.weight for c in m.classifier]), torch.stack([c.fc.bias for c in m.classifier]))
    head = (m.Slide_classifier.fc.weight, m.Slide_classifier.fc.bias)
    import acmil_b200._lib as L_
    capacity = L_.record_floats(5, 128, 10) * 4 * len(sizes)
    bufs = [torch.zeros((256 + 2 * ranks * capacity) // 4, device="cuda") for _ in range(ranks)]
    xch = [_LocalExchange(ranks, r, bufs, capacity) for r in range(ranks)]
    off = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    for step in range(3):
        x = torch.randn(sum(sizes), 384, generator=g).cuda()
        rand = torch.rand(len(sizes), 5, 10, generator=g).cuda()
        kw = dict(n_masked=10, keep=[6, 6], rand=rand, branch_w=branch[0], branch_b=branch[1], head_w=head[0], head_b=head[1],
                  slide_head=True)
        with torch.no_grad():
            full = op.run(packed, x, off, **kw)
            ctxs, locs = [], []
            for r in range(ranks):      # all pushes first (one stream: a finish would wait for flags nobody raised yet)
                b = [[n * q // ranks for q in range(ranks + 1)] for n in sizes]
                xs = torch.cat([x[off[s] + b[s][r]: off[s] + b[s][r + 1]] for s in range(len(sizes))])
                lo = np.concatenate([[0], np.cumsum([b[s][r + 1] - b[s][r] for s in range(len(sizes))])]).tolist()
                _, ctx = op.partial(packed, xs, lo, n_masked=10, shard_begin=[b[s][r] for s in range(len(sizes))], exchange=xch[r])
                ctxs.append(ctx)
                locs.append((b, lo))
            fkw = {k: v for k, v in kw.items() if k != "n_masked"}
            outs = [op.finish(ctxs[r], None, ranks, **fkw) for r in range(ranks)]
        for r, o in enumerate(outs):
            assert torch.equal(o.masked_idx.sort(-1).values, full.masked_idx.sort(-1).values), (step, r)
            close(o.sub, full.sub.cpu().numpy(), rtol=2e-5, atol=1e-6)
            close(o.slide, full.slide.cpu().numpy(), rtol=2e-5, atol=1e-6)
            b, lo = locs[r]
            for s in range(len(sizes)):
                close(o.scores[:, lo[s]:lo[s + 1]], full.scores[:, off[s] + b[s][r]: off[s] + b[s][r + 1]].cpu().numpy(),
                      rtol=1e-6, atol=1e-7)
        assert all(int(xc.state[0]) == step + 1 for xc in xch)      # every rank's epoch advanced once per step


def _grad_close(name, got, ref, strict=True, floor=0.0):
    """strict: every entry within 2e-3 of the tensor's largest entry.  Not strict (the two sides computed the front layer
    independently): a pre-activation within rounding distance of the ReLU's kink (|z| ~ 1e-5) may fall on the other side,
    which moves single entries of dz1 by their whole value and with them one row of dx / dW1 -- both answers are valid
    subgradients -- so only the tensor-level l2 error is bounded (5e-3)."""
    d = (got - ref).abs()
    if strict:
        # (floor: tensors whose true gradient nearly cancels -- the score bias under a softmax, everything softmax-related in a
        # one-row bag -- are measured against the largest gradient of the step instead of their own rounding noise)
        err = float(d.max()) / max(float(ref.abs().max()), floor, 1e-30)
        assert err < 2e-3, (name, err)
    else:
        l2 = float(d.norm()) / (float(ref.norm()) + 1e-30)
        assert l2 < 5e-3, (name, l2)


@pytest.mark.parametrize("case", [
    dict(d_in=384, d_inner=128, K=5, n=2999, masked=True),                                   # ACMIL_GA, n % 4 != 0
    dict(d_in=384, d_inner=128, K=1, n=1024),                                                # ABMIL
    dict(d_in=512, d_inner=256, K=3, n=1500, front_bias=True),                               # CLAM_MB-like (Linear + bias)
    dict(d_in=512, d_inner=256, K=1, n=801, front_bias=True, gated=False),                   # CLAM_SB(gate=False)
    dict(d_in=1024, d_inner=512, K=1, n=700, front_bias=True, act_a="gelu", biases=False),   # attmil.AttentionGated
    dict(d_in=1024, d_inner=512, K=2, n=640, act_a="relu"),
    dict(d_in=384, d_inner=128, K=5, n=7, masked=True),                                      # fewer rows than n_masked_patch
    dict(d_in=384, d_inner=128, K=3, n=1),                                                   # a one-patch bag
])
def test_pool_backward_kernels_match_autograd(case):
    """acmil_b200.gp_backward (tcgen05 GEMMs + csrc/gp_bwd.cu) against torch autograd of the same graph written with
    torch ops in fp32 (TF32 off), for every gate flavour of the family; gradients w.r.t. all weights and x."""
    import torch.nn.functional as F
    from acmil_b200 import gp_backward as B
    from acmil_b200.gated_pool import GatedPool, GatedPoolSpec
    torch.backends.cuda.matmul.allow_tf32 = False
    d_in, Li, K, n = case["d_in"], case["d_inner"], case["K"], case["n"]
    gated, act_a, fb, biases = case.get("gated", True), case.get("act_a", "tanh"), case.get("front_bias", False), case.get("biases", True)
    spec = GatedPoolSpec(d_in=d_in, d_inner=Li, n_branch=K, front_bias=fb, act_a=act_a, gated=gated, gate_bias=biases,
                         score_bias=biases)
    assert B.supported(spec)
    g = torch.Generator().manual_seed(1000 + n)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()      # noqa: E731
    w = dict(w1=rnd(Li, d_in, scale=d_in ** -0.5), b1=rnd(Li, scale=0.1) if fb else None, wv=rnd(128, Li, scale=Li ** -0.5),
             bv=rnd(128, scale=0.1) if biases else None, wu=rnd(128, Li, scale=Li ** -0.5) if gated else None,
             bu=rnd(128, scale=0.1) if (gated and biases) else None, ww=rnd(K, 128, scale=0.3),
             bw=rnd(K, scale=0.1) if biases else None)
    x = rnd(n, d_in)
    op = GatedPool(spec)
    packed = op.pack(w["w1"], w["b1"], w["wv"], w["bv"], w["wu"], w["bu"], w["ww"], w["bw"])
    n_masked, keep, rand = (10, [int(min(10, n) * 0.6)], torch.rand(1, K, 10, generator=g).cuda()) if case.get("masked") else (0, [0], None)
    res = op.run(packed, x, [0, n], n_masked=n_masked, keep=keep, rand=rand)
    g_afeat, g_bag, g_scores = rnd(K, Li), rnd(1, Li), rnd(K, n, scale=1e-3)
    dbg = {}
    got = B.pool_backward(spec, x, {k: v for k, v in w.items() if v is not None}, res.scores, res.lse_m[0], res.lse_l[0],
                          res.afeat[0], g_afeat, g_bag, g_scores, need_dx=True, _debug=dbg)
    # the same graph with torch ops.  The ReLU of the front layer uses the kernels' own sign pattern: a pre-activation
    # within rounding distance of 0 (a handful of the n * d_inner entries) may otherwise fall on either side of the kink
    leaves = {k: v.clone().requires_grad_(True) for k, v in w.items() if v is not None}
    xr = x.clone().requires_grad_(True)
    z1 = F.linear(xr, leaves["w1"], leaves.get("b1"))
    flipped = z1.detach()[(dbg["h"] > 0) != (z1.detach() > 0)]
    assert flipped.numel() <= max(3, int(1e-5 * z1.numel())) and (flipped.numel() == 0 or float(flipped.abs().max()) < 1e-4)
    close(dbg["h"], F.relu(z1).detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    h = z1 * (dbg["h"] > 0)
    zv = F.linear(h, leaves["wv"], leaves.get("bv"))
    a = torch.tanh(zv) if act_a == "tanh" else (F.relu(zv) if act_a == "relu" else F.gelu(zv))
    if gated:
        a = a * torch.sigmoid(F.linear(h, leaves["wu"], leaves.get("bu")))
    s = F.linear(a, leaves["ww"], leaves.get("bw")).t()
    s = s.masked_fill(res.scores == -1e9, -1e9)
    close(res.scores, s.detach().cpu().numpy(), rtol=1e-3, atol=1e-5)
    af = torch.softmax(s, 1) @ h
    outs = [af, af.mean(0, keepdim=True), s]
    names = list(leaves)
    ref = torch.autograd.grad(outs, [xr] + [leaves[k] for k in names], [g_afeat, g_bag, g_scores])
    floor = 1e-2 * max(float(r.abs().max()) for r in ref[1:])
    for k, r in zip(["x"] + names, ref):
        assert got[k].shape == r.shape, k
        _grad_close(k, got[k], r, floor=floor if k != "x" else 0.0)


def test_acmil_training_step_kernel_vs_torch_backward(monkeypatch):
    """A full training step through the module (forward kernels + kernel backward + AdamW) gives the same updated weights
    as the same step with the torch-op backward (ACMIL_POOL_BACKWARD=torch)."""
    import torch.nn.functional as F
    x = torch.randn(1, 4000, 384, generator=torch.Generator().manual_seed(8)).cuda()
    y = torch.tensor([1], device="cuda")
    results = []
    for mode in ("kernel", "torch"):
        monkeypatch.setenv("ACMIL_POOL_BACKWARD", mode)
        m = _random_acmil(33).cuda().train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
        torch.manual_seed(5)
        sub, slide, a = m(x)
        p = torch.softmax(a, dim=-1)
        d = sum(torch.cosine_similarity(p[:, i], p[:, j], dim=-1).mean() for i in range(5) for j in range(i + 1, 5)) / 10
        (F.cross_entropy(sub, y.repeat_interleave(5)) + F.cross_entropy(slide, y) + d).backward()
        results.append({k: v.grad.clone() for k, v in m.named_parameters()})
        opt.step()
    # (the score bias' gradient is zero up to rounding -- softmax is shift-invariant -- so errors are measured against the
    # largest gradient of the step where a tensor's own norm is negligible)
    top = max(float(v.norm()) for v in results[1].values())
    for k in results[0]:
        ref = results[1][k]
        l2 = float((results[0][k] - ref).norm()) / max(float(ref.norm()), 1e-4 * top)
        assert l2 < 5e-3, (k, l2)


@pytest.mark.parametrize("K,n", [(5, 50000), (2, 777), (8, 4096)])
def test_diversity_loss_matches_the_training_script(K, n):
    """acmil_b200.losses.diversity_loss against Step3_WSI_classification_ACMIL.py:208-214 spelled with torch ops: value and
    gradient w.r.t. the raw scores, with masked (-1e9) entries."""
    from acmil_b200.losses import diversity_loss
    g = torch.Generator().manual_seed(K * 1000 + n)
    a = (torch.randn(1, K, n, generator=g) * 2).cuda()
    a[0, :, 5] = -1e9
    a[0, 0, 11] = -1e9
    a1 = a.clone().requires_grad_(True)
    p = torch.softmax(a1, dim=-1)
    ref = sum(torch.cosine_similarity(p[:, i], p[:, j], dim=-1).mean() / (K * (K - 1) / 2) for i in range(K) for j in range(i + 1, K))
    (3.0 * ref).backward()
    a2 = a.clone().requires_grad_(True)
    mine = diversity_loss(a2)
    (3.0 * mine).backward()
    assert abs(float(mine) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    err = float((a2.grad - a1.grad).abs().max()) / (float(a1.grad.abs().max()) + 1e-30)
    assert err < 1e-3, err
    assert float(a2.grad[0, :, 5].abs().max()) == 0.0
    assert float(diversity_loss(a[:, :1])) == 0.0      # one branch: the reference's double loop is empty


@pytest.mark.parametrize("impl_name", ["umma", "ffma"])
@pytest.mark.parametrize("train", [False, True])
def test_fp16_rows_equal_widened_rows(impl_name, train):
    """x_f16: the kernels read fp16 features (the reference's H5 storage dtype, Step2_feature_extract.py:165) directly.
    Widening fp16 -> fp32 is exact, so every output must equal the one computed from the widened rows: masks identical,
    scores / logits to rounding (the tcgen05 kernel sums hi products only: same values, other order)."""
    import acmil_b200._lib as L
    m = _random_acmil(41).cuda()
    m._op.impl = L.IMPL_UMMA if impl_name == "umma" else L.IMPL_FFMA
    g = torch.Generator().manual_seed(6)
    sizes = [3000, 1, 4097]
    x16 = torch.randn(sum(sizes), 384, generator=g).half().cuda()
    off = [0, 3000, 3001, 3001 + 4097]
    rand = torch.rand(3, 5, 10, generator=g).cuda()
    m.train(train)
    with torch.no_grad():
        a = m.forward_bags(x16, off, rand=rand if train else None)
        b = m.forward_bags(x16.float(), off, rand=rand if train else None)
    for u, v in zip(a, b):
        assert np.array_equal((u == -1e9).cpu().numpy(), (v == -1e9).cpu().numpy())
        close(u, v.cpu().numpy(), rtol=1e-5, atol=2e-6)
    # through the single-bag module call too (autograd path in train mode is covered by the backward tests)
    with torch.no_grad():
        torch.manual_seed(3)
        s16 = m(x16[None, :3000])
        torch.manual_seed(3)
        s32 = m(x16[None, :3000].float())
    for u, v in zip(s16, s32):
        close(u, v.cpu().numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("case", [
    dict(d_in=512, d_inner=256, K=3),                                                  # ResNet18 / CLIP features (Step3 table)
    dict(d_in=1024, d_inner=512, K=1, front_bias=True, front_act="gelu", gated=False),   # attmil.DAttention
    dict(d_in=768, d_inner=384, K=5),
])
def test_front_projection_on_the_gemm_engine_matches_the_fused_fp32_kernel(case):
    """Shapes outside the fused tcgen05 kernel: in eval mode the front projection runs on the tcgen05 GEMM engine (3xTF32)
    and the exact FFMA kernel pools h; the result must match the all-FFMA (exact fp32) run of the same head."""
    import acmil_b200._lib as L
    from acmil_b200.gated_pool import GatedPool, GatedPoolSpec
    if case["d_inner"] % 128:
        pytest.skip("d_inner must be a multiple of 128 for the pool kernels")
    d_in, Li, K = case["d_in"], case["d_inner"], case["K"]
    gated, fb, fa = case.get("gated", True), case.get("front_bias", False), case.get("front_act", "relu")
    spec = GatedPoolSpec(d_in=d_in, d_inner=Li, n_branch=K, front_bias=fb, front_act=fa, gated=gated)
    g = torch.Generator().manual_seed(d_in)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda()      # noqa: E731
    w = (rnd(Li, d_in, scale=d_in ** -0.5), rnd(Li, scale=0.1) if fb else None, rnd(128, Li, scale=Li ** -0.5), rnd(128, scale=0.1),
         rnd(128, Li, scale=Li ** -0.5) if gated else None, rnd(128, scale=0.1) if gated else None, rnd(K, 128, scale=0.3),
         rnd(K, scale=0.1))
    n = [5000, 4500]
    x = rnd(sum(n), d_in)
    off = [0, n[0], sum(n)]
    auto, exact = GatedPool(spec, L.IMPL_AUTO), GatedPool(spec, L.IMPL_FFMA)
    ra = auto.run(auto.pack(*w), x, off)
    assert auto._tail is not None                       # the composed path was taken
    re = exact.run(exact.pack(*w), x, off)
    # (3xTF32 front projection over K up to 1024 + the A&S erf of its GELU epilogue: a few 1e-5 absolute on scores of O(0.1))
    close(ra.scores, re.scores.cpu().numpy(), rtol=1e-3, atol=5e-5)
    close(ra.afeat, re.afeat.cpu().numpy(), rtol=1e-3, atol=5e-5)
    close(ra.bag_feat, re.bag_feat.cpu().numpy(), rtol=1e-3, atol=5e-5)
    # masking keeps the exact kernel end to end (mask indices are promised bit-exact)
    rm = auto.run(auto.pack(*w), x, off, n_masked=10, keep=[6, 6], rand=torch.rand(2, K, 10, generator=g).cuda())
    rx = exact.run(exact.pack(*w), x, off, n_masked=10, keep=[6, 6], rand=torch.rand(2, K, 10, generator=torch.Generator().manual_seed(1)).cuda())
    assert torch.equal(rm.topk_idx, rx.topk_idx)


@pytest.mark.parametrize("name", ["acmil_ga_k5_n1024", "acmil_ga_k3_d512_n2000", "acmil_ga_k8_n4099"])
def test_public_train_forward_and_masked_forward_feature_golden(name, monkeypatch):
    """The PUBLIC calls -- ``forward`` in train mode and ``forward_feature(x, use_attention_mask=True)`` (transformer.py:305-352)
    -- with the reference's own uniform draw replayed through torch.rand: masked positions bit-exact, outputs within 1e-3."""
    import acmil_b200.heads as Hd
    w, g = load_golden(name)
    x = golden_x(g).cuda()
    m = make_acmil(w, g, 0)
    draw = torch.from_numpy(g["train_rand"]).cuda()
    calls = []

    def fake_rand(*shape, **kw):
        assert tuple(shape) == tuple(draw.shape), (shape, draw.shape)      # the reference's call shape (transformer.py:316)
        calls.append(1)
        return draw

    monkeypatch.setattr(Hd.torch, "rand", fake_rand)
    m.train()
    with torch.no_grad():
        sub, slide, a = m(x)
        m.eval()                                                            # forward_feature masks by its argument, not by mode
        feat = m.forward_feature(x, use_attention_mask=True)
        feat_plain = m.forward_feature(x)
    monkeypatch.undo()
    assert len(calls) == 2
    a = a.cpu().numpy()
    assert np.array_equal(a == -1e9, g["train_A"] == -1e9)
    both(a, g["train_A"])
    both(sub, g["train_sub"])
    both(slide, g["train_slide"])
    both(feat, g["train_feat"])
    both(feat_plain, g["eval_feat"])
