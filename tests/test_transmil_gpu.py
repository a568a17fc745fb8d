"""GPU parity tests of the TransMIL / Nystrom path (through the C-ABI): the tcgen05 GEMM against fp64 matmul, the
small kernels against torch, and the modules against golden vectors made by the reference itself and against
the numpy oracle.  Tolerance per BASELINE.json north_star: logits within 1e-3 relative."""
import numpy as np
import pytest
import torch

from conftest import golden_names, golden_x, load_golden

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def load_into(module, w):
    module.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return module.to(dev()).eval()


@pytest.mark.parametrize("m,n,k,batch", [(128, 128, 32, 1), (300, 200, 100, 1), (1, 2, 64, 1), (257, 64, 512, 3),
                                          (64, 256, 1000, 2), (1000, 48, 36, 1), (130, 130, 4, 2)])
def test_gemm_nt_matches_fp64(m, n, k, batch):
    from acmil_b200.transmil import gemm_nt
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(batch, m, k, generator=g)
    b = torch.randn(batch, n, k, generator=g)
    ref = (a.double() @ b.double().transpose(1, 2))
    scale = float(ref.abs().max())
    out = gemm_nt(a.to(dev()), b.to(dev())).cpu()
    err = float((out.double() - ref).abs().max()) / scale
    assert err < 2e-5, err                     # 3xTF32 (tensor-core fp32 accumulation truncates: ~k * 2^-24)
    out1 = gemm_nt(a.to(dev()), b.to(dev()), precise=False).cpu()
    err1 = float((out1.double() - ref).abs().max()) / scale
    assert err1 < 5e-3 and (k < 16 or err1 > err)      # plain TF32 is the coarse mode


def test_gemm_nt_epilogue_terms_and_layouts():
    from acmil_b200.transmil import gemm_nt
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(2, 200, 72, generator=g), torch.randn(200, 72, generator=g)      # shared B
    bias, add = torch.randn(200, generator=g), torch.randn(2, 200, 200, generator=g)
    ref = torch.relu(-0.5 * (a.double() @ b.double().T) + 3.0 * torch.eye(200, dtype=torch.float64) + bias.double()
                     + 0.25 * add.double())
    out_t = torch.empty(2, 200, 200, device=dev())
    out = gemm_nt(a.to(dev()), b.to(dev()), bias=bias.to(dev()), addend=add.to(dev()), alpha=-0.5, beta=0.25, diag=3.0,
                  relu=True, out_t=out_t)
    assert float((out.cpu().double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    assert torch.equal(out_t, out.transpose(1, 2))
    # K split (long reductions): same result up to summation order
    a2, b2 = torch.randn(3, 64, 5000, generator=g), torch.randn(3, 96, 5000, generator=g)
    ref2 = a2.double() @ b2.double().transpose(1, 2)
    out2 = gemm_nt(a2.to(dev()), b2.to(dev()), k_split=7).cpu()
    assert float((out2.double() - ref2).abs().max()) / float(ref2.abs().max()) < 2e-5
    # strided output view (rows of a larger buffer)
    buf = torch.zeros(2, 210, 200, device=dev())
    gemm_nt(a.to(dev()), b.to(dev()), out=buf[:, 5:205])
    ref3 = a.double() @ b.double().T
    assert float((buf[:, 5:205].cpu().double() - ref3).abs().max()) < 2e-5 * float(ref3.abs().max())
    assert float(buf[:, :5].abs().max()) == 0 and float(buf[:, 205:].abs().max()) == 0


@pytest.mark.parametrize("m,n,k,wscale", [(300, 200, 384, 0.02), (1000, 384, 1536, 0.02), (257, 70, 40, 1.0), (128, 64, 100, 50.0),
                                           (513, 130, 64, 1e-4), (50, 8, 24, 0.3), (2000, 1536, 384, 0.02),
                                           # CTA-pair tilings (>= 74 pair tiles): 256-wide with an N tail inside the second CTA's
                                           # half of B and a last pair whose second half lies wholly past M; 128-wide; no tails
                                           (10084, 400, 200, 0.02), (10084, 200, 96, 1.0), (9728, 512, 512, 0.05)])
def test_gemm_fp16_split_weights_matches_fp64(m, n, k, wscale):
    """precise = 2: B from a pre-split fp16 hi / lo image of the weight (acmil_gemm_split_b), A split in the kernel.  The
    K tails (40, 100: the second 32-column A box of the last chunk is partly / wholly out of range), N / M tails and the
    power-of-two weight scaling (weights of 1e-4 .. 50) all against fp64; accuracy must be that of the 3xTF32 mode."""
    from acmil_b200.transmil import SplitImage, gemm_nt
    g = torch.Generator().manual_seed(m + 3 * n + 7 * k)
    a = torch.randn(m, k, generator=g)
    a[::7] *= 30.0                                   # rows of different magnitude: no per-row scaling is assumed
    a[1::5] *= 1e-3
    w = torch.randn(n, k, generator=g) * wscale
    bias = torch.randn(n, generator=g)
    ref = a.double() @ w.double().T
    # error scale of each dot product: the fp16 lo part of a has an absolute floor of 2^-25 (elements below 0.25)
    bound = ((a.double().abs() + 0.125) @ w.double().abs().T)
    img = SplitImage(w.to(dev()))
    out = gemm_nt(a.to(dev()), w.to(dev()), b_split=img).cpu().double()
    err = float(((out - ref).abs() / bound.clamp(min=1e-30)).max())
    out1 = gemm_nt(a.to(dev()), w.to(dev())).cpu().double()
    err1 = float(((out1 - ref).abs() / bound.clamp(min=1e-30)).max())
    # both splits are exact to ~2^-22 per product; what is left is the tensor core's truncating fp32 accumulation
    # (measured on B200: 7.6e-7 .. 8.5e-7 of sum |a||w| for the fp16 split), the same for either kind
    assert err < 3e-6 and err < 4 * err1 + 5e-7, (err, err1)
    # epilogue terms ride on the same code as the TF32 kernels
    add = torch.randn(m, n, generator=g)
    out2 = gemm_nt(a.to(dev()), w.to(dev()), b_split=img, bias=bias.to(dev()), addend=add.to(dev()), alpha=0.5, beta=2.0,
                   relu=True).cpu().double()
    ref2 = torch.relu(0.5 * ref + bias.double() + 2.0 * add.double())
    assert float(((out2 - ref2).abs() / (bound + 1.0)).max()) < 1e-6


def test_gemm_fp16_split_row_ranges_batches_and_errors():
    from acmil_b200 import _lib as L
    from acmil_b200.transmil import SplitImage, gemm_nt
    g = torch.Generator().manual_seed(11)
    w = torch.randn(3 * 96, 128, generator=g) * 0.05                      # a to_qkv-like weight: three row blocks
    a = torch.randn(4, 150, 128, generator=g)                             # batched A, shared weight
    img = SplitImage(w.to(dev()))
    for blk in range(3):
        wb = w[blk * 96:(blk + 1) * 96]
        out = gemm_nt(a.to(dev()), wb.to(dev()), b_split=img, b_split_row0=blk * 96).cpu().double()
        ref = a.double() @ wb.double().T
        assert float((out - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    # batched A against a shared pre-split weight on the CTA-pair kernel (the ViT patch embedding's shape: 196 rows per image)
    a2 = torch.randn(40, 196, 128, generator=g)
    out = gemm_nt(a2.to(dev()), w.to(dev()), b_split=img).cpu().double()
    ref = a2.double() @ w.double().T
    assert float((out - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    # an all-zero weight (scale falls back to 1) and a strided weight view
    z = torch.zeros(16, 64)
    assert float(gemm_nt(a[0, :, :64].contiguous().to(dev()), z.to(dev()), b_split=SplitImage(z.to(dev()))).abs().max()) == 0.0
    wv = torch.randn(40, 200, generator=g).to(dev())[:, :128]
    out = gemm_nt(a[0].to(dev()), wv, b_split=SplitImage(wv)).cpu().double()
    ref = a[0].double() @ wv.cpu().double().T
    assert float((out - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    with pytest.raises(ValueError):
        gemm_nt(a.to(dev()), w[:96].to(dev()), b_split=img, b_split_row0=250)
    with pytest.raises(ValueError):
        gemm_nt(a[..., :64].contiguous().to(dev()), w[:96, :64].contiguous().to(dev()), b_split=img)
    with pytest.raises(L.AcmilError):      # a pre-split B cannot be k-split
        gemm_nt(a.to(dev()), w[:96].to(dev()), b_split=img, k_split=2)


@pytest.mark.parametrize("pair", ["0", "2"])
def test_gemm_fp16_split_other_tilings_in_a_subprocess(pair):
    """The tiling of the fp16-split products is chosen once per process (ACMIL_GEMM_PAIR; default 1 = 128-wide CTA pairs, what
    every other test in this file runs): the one-CTA-per-tile kernel (0) and the 256-wide pair tiles (2) get the same shapes
    in a fresh process each."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from acmil_b200.transmil import SplitImage, gemm_nt\n"
        "torch.manual_seed(0)\n"
        "for m, n, k in ((10084, 400, 200), (9728, 512, 512), (10084, 200, 96), (300, 200, 384)):\n"
        "    a = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda') * 0.05\n"
        "    out = gemm_nt(a, w, b_split=SplitImage(w)).double()\n"
        "    ref = a.double() @ w.double().T\n"
        "    err = float((out - ref).abs().max() / ref.abs().max())\n"
        "    assert err < 1e-5, (m, n, k, err)\n"
        "print('tilings ok')\n")
    env = dict(os.environ, ACMIL_GEMM_PAIR=pair)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "tilings ok" in r.stdout, r.stdout + r.stderr


def test_modules_agree_between_the_fp16_split_and_the_tf32_split(monkeypatch):
    """The weight products of TransMIL run on the fp16-split kernel by default; ACMIL_GEMM_SPLIT=tf32 keeps every product
    on 3xTF32.  Both are fp32-faithful: the logits agree to 1e-4 relative (and each is checked against the reference's
    fixtures elsewhere in this file under the default)."""
    from acmil_b200 import Struct, TransMIL
    import acmil_b200.transmil as T
    torch.manual_seed(3)
    m = TransMIL(Struct(D_feat=384, D_inner=128, n_class=3)).to(dev()).eval()
    x = torch.randn(1, 1500, 384, device=dev())
    with torch.no_grad():
        assert T.gemm_mode() == 2
        y2 = m(x)
        monkeypatch.setenv("ACMIL_GEMM_SPLIT", "tf32")
        assert T.gemm_mode() == 1
        y1 = m(x)
    assert float((y1 - y2).abs().max()) <= 1e-4 * max(1.0, float(y1.abs().max()))


@pytest.mark.parametrize("m,n,d,batch", [(197, 200, 64, 6), (300, 256, 64, 2), (64, 40, 32, 3), (130, 1000, 16, 1)])
def test_gemm_chunked_softmax_across_two_products(m, n, d, batch):
    """softmax(q k^T) v without a softmax pass: the first product's epilogue stores exp(s - chunk max) per 32-column chunk
    plus (max, sum) per chunk, the second product's operand converter rescales every chunk of a row by
    exp(max_c - row max) / row sum.  Against torch.softmax in fp64, including N / K tails inside a chunk and rows whose
    chunks differ by e^20 in magnitude."""
    from acmil_b200.transmil import gemm_nt
    g = torch.Generator().manual_seed(m + n)
    q, k, v = (torch.randn(batch, r, c, generator=g) for r, c in ((m, d), (n, d), (n, d)))
    q[:, ::3] *= 4.0                                  # peaked rows
    ref_p = torch.softmax(0.7 * q.double() @ k.double().transpose(1, 2), dim=-1)
    ref = ref_p @ v.double()
    nch = (n + 31) // 32
    stats = torch.empty(batch, m, nch, 2, device=dev())
    e = gemm_nt(q.to(dev()), k.to(dev()), alpha=0.7, stats_out=stats)
    assert float(e.max()) <= 1.0 and float(e.min()) >= 0.0
    out = gemm_nt(e, v.transpose(1, 2).contiguous().to(dev()), stats_in=stats).cpu().double()
    assert float((out - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
    # the pieces themselves: e * exp(max_c - M) / sum == softmax
    st = stats.cpu().double()
    M = st[..., 0].max(-1, keepdim=True).values
    L = (st[..., 1] * torch.exp(st[..., 0] - M)).sum(-1, keepdim=True)
    f = (torch.exp(st[..., 0] - M) / L).repeat_interleave(32, dim=-1)[..., :n]
    assert float((e.cpu().double() * f - ref_p).abs().max()) < 1e-5      # __expf on arguments up to ~60: |x| 2^-23 relative


def test_gemm_gelu_epilogue_is_the_exact_erf_gelu_within_6e7():
    """The fc1 epilogue of the ViT MLP evaluates nn.GELU (exact-erf, timm's default act_layer) with the
    Abramowitz-Stegun 7.1.26 erf on packed fp32 -- a deliberate deviation from libm's erff.  Bound it by itself, over
    [-10, 10] and around 0: |gelu_kernel(x) - gelu_fp64(x)| <= 6e-7 * max(1, |x|) (A&S: |erf error| <= 1.5e-7, its
    exponential on ex2.approx, fp32 rounding of the polynomial and of the result; measured on B200: 4.8e-7); the GEMM in
    front is x @ I^T, exact for the hi/lo split.  3e-4 of the fp16 store's own rounding (Step2_feature_extract.py:165)."""
    from acmil_b200.transmil import gemm_nt
    xs = torch.cat([torch.linspace(-10, 10, 128 * 511), torch.linspace(-1e-3, 1e-3, 128)]).reshape(-1, 128).contiguous()
    eye = torch.eye(128)
    out = gemm_nt(xs.to(dev()), eye.to(dev()), gelu=True).cpu().double()
    ref = torch.nn.functional.gelu(xs.double())
    err = (out - ref).abs() / xs.double().abs().clamp(min=1.0)
    assert float(err.max()) <= 6e-7, float(err.max())
    # and the fp32 libm GELU torch itself computes on the GPU is no closer to fp64 than twice that
    ref32 = torch.nn.functional.gelu(xs.to(dev())).cpu().double()
    assert float(((out - ref32).abs() / xs.double().abs().clamp(min=1.0)).max()) <= 1e-6


def test_gemm_rejects_bad_arguments():
    from acmil_b200 import _lib as L
    from acmil_b200.transmil import gemm_nt
    with pytest.raises(RuntimeError):
        gemm_nt(torch.randn(4, 8), torch.randn(4, 8).to(dev()))
    with pytest.raises(ValueError):
        gemm_nt(torch.randn(4, 8).to(dev()), torch.randn(4, 12).to(dev()))
    with pytest.raises(L.AcmilError):      # lda not a multiple of 4 floats
        gemm_nt(torch.randn(4, 6).to(dev()), torch.randn(4, 6).to(dev()))


def test_layernorm_rows_matches_torch():
    from acmil_b200.transmil import layernorm_rows
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1000, 512, generator=g) * 3 + 1
    w, b = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (512,), w.double(), b.double(), 1e-5)
    out = layernorm_rows(x.to(dev()), w.to(dev()), b.to(dev()), 1e-5).cpu()
    assert float((out.double() - ref).abs().max()) < 1e-5


@pytest.mark.parametrize("name", golden_names("ppeg_"))
def test_ppeg_matches_reference_golden(name):
    from acmil_b200.transmil import PPEG
    w, meta = load_golden(name)
    dim, gh, gw = (int(v) for v in meta["meta_cfg"])
    mod = load_into(PPEG(dim=dim), w)
    with torch.no_grad():
        y = mod(golden_x(meta).to(dev()), gh, gw).cpu().numpy()
    np.testing.assert_allclose(y, meta["out"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", golden_names("nystrom_"))
def test_nystrom_attention_matches_reference_golden(name):
    from acmil_b200.transmil import NystromAttention
    w, meta = load_golden(name)
    dim, dim_head, heads, m, ks, residual, iters = (int(v) for v in meta["meta_cfg"])
    mod = load_into(NystromAttention(dim=dim, dim_head=dim_head, heads=heads, num_landmarks=m, pinv_iterations=iters,
                                     residual=bool(residual), residual_conv_kernel=ks, dropout=0.1), w)
    with torch.no_grad():
        y = mod(golden_x(meta).to(dev())).cpu().numpy()
    assert y.shape == meta["out"].shape
    np.testing.assert_allclose(y, meta["out"], rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize("name", golden_names("translayer_"))
def test_trans_layer_matches_reference_golden(name):
    from acmil_b200.transmil import TransLayer
    w, meta = load_golden(name)
    mod = load_into(TransLayer(dim=int(meta["meta_cfg"][0])), w)
    x = golden_x(meta).to(dev())
    with torch.no_grad():
        y = mod(x).cpu().numpy()
        y3 = mod(x, n_out=3).cpu().numpy()
    np.testing.assert_allclose(y, meta["out"], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(y3, meta["out"][:, :3], rtol=1e-3, atol=2e-5)      # the rows-subset path of layer 2


@pytest.mark.parametrize("name", golden_names("transmil_"))
def test_transmil_matches_reference_golden(name):
    from acmil_b200 import Struct
    from acmil_b200.transmil import TransMIL
    w, meta = load_golden(name)
    d_feat, d_inner, n_class = (int(v) for v in meta["meta_cfg"])
    mod = load_into(TransMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class)), w)
    with torch.no_grad():
        y = mod(golden_x(meta).to(dev())).cpu().numpy()
    np.testing.assert_allclose(y, meta["out"], rtol=1e-3, atol=1e-4)      # north_star: logits within 1e-3


def test_transmil_dim512_vs_oracle_and_plain_tf32_mode():
    """BASELINE config 3 dims (D_feat 512, D_inner 512 -> 256 landmarks, 8 x 64 heads) at an oracle-sized n."""
    from acmil_b200 import Struct
    from acmil_b200.transmil import TransMIL
    from oracle import transmil as O
    torch.manual_seed(3)
    mod = TransMIL(Struct(D_feat=512, D_inner=512, n_class=2)).eval()
    x = torch.randn(1, 3000, 512, generator=torch.Generator().manual_seed(4))
    w = {k: v.numpy() for k, v in mod.state_dict().items()}
    ref = O.transmil_forward(w, x.numpy())
    mod = mod.to(dev())
    with torch.no_grad():
        y = mod(x.to(dev())).cpu().numpy()
        np.testing.assert_allclose(y, ref, rtol=1e-3, atol=1e-4)
        for layer in (mod.layer1, mod.layer2):
            layer.attn.precise = False
        y1 = mod(x.to(dev())).cpu().numpy()
    assert np.isfinite(y1).all() and np.abs(y1 - ref).max() < 0.1      # coarse mode: sane, not parity-grade


def _seeded_transmil(meta):
    """acmil_b200.TransMIL built under the fixture's model seed; its weights must be the ones the reference drew."""
    import hashlib
    from acmil_b200 import Struct
    from acmil_b200.transmil import TransMIL
    d_feat, d_inner, n_class = (int(v) for v in meta["meta_cfg"])
    torch.manual_seed(int(meta["meta_model_seed"]))
    mod = TransMIL(Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class)).eval()
    h = hashlib.sha256()
    for k, v in mod.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    assert h.hexdigest() == str(meta["meta_w_sha"]), "seeded construction differs from the reference's"
    return mod


@pytest.mark.parametrize("name", golden_names("seeded_transmil_"))
def test_transmil_seeded_reference_fixtures(name):
    """SURVEY KAT4 (dim 512, n = 1000) and BASELINE.json configs[2] AT ITS OWN SIZE (N = 50 000, dim 512, 256 landmarks):
    logits of the reference itself (tests/golden/make_golden_transmil.py), weights by seed.  At N = 50k this is the only
    place where the one-CTA-per-50k-row softmax, the K-split attn3 . v and the 255 front-pad rows are checked."""
    _, meta = load_golden(name)
    mod = _seeded_transmil(meta).to(dev())
    x = golden_x(meta).to(dev())
    with torch.no_grad():
        y0 = mod(x)
        y1 = mod(x)
    assert torch.equal(y0, y1)                                                        # deterministic
    np.testing.assert_allclose(y0.cpu().numpy(), meta["out"], rtol=1e-3, atol=1e-4)   # north_star: logits within 1e-3


def test_forward_only_and_cuda_only_errors():
    from acmil_b200 import Struct
    from acmil_b200.transmil import NystromAttention, TransMIL
    mod = TransMIL(Struct(D_feat=32, D_inner=64, n_class=2)).to(dev())
    with pytest.raises(NotImplementedError):
        mod(torch.randn(1, 40, 32, device=dev()))              # parameters require grad, grad mode on
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            mod(torch.randn(1, 40, 32))                        # CPU tensor
    att = NystromAttention(64, dim_head=8, heads=8, num_landmarks=32).to(dev())
    with pytest.raises(NotImplementedError), torch.no_grad():
        att(torch.randn(1, 40, 64, device=dev()), mask=torch.ones(1, 40, dtype=torch.bool, device=dev()))


def _sharded_logits(mod, x, world):
    """TransMIL over `world` ranks played by threads on this GPU (ThreadComm): the per-rank code of the multi-GPU path."""
    from acmil_b200.transmil_sharded import ShardPlan, run_threads, transmil_forward_sharded
    n = x.shape[1]
    plan = ShardPlan(n, world, mod.layer1.attn.num_landmarks)

    def rank_fn(rank, comm):
        rows = torch.from_numpy(plan.patch_rows(rank)).to(x.device)
        return transmil_forward_sharded(mod, x[0].index_select(0, rows), n, rank, comm)

    return run_threads(world, rank_fn)


@pytest.mark.parametrize("d_inner,n,world", [(64, 300, 2), (64, 1000, 4), (128, 1000, 8), (64, 50, 1)])
def test_transmil_sharded_equals_unsharded(d_inner, n, world):
    """Sequence-parallel TransMIL (landmark-aligned shards, log-sum-exp merge of attn3 v, pinv heads split over the ranks,
    conv / PPEG halos) against the single-GPU forward of the same module: every rank returns the same logits."""
    from acmil_b200 import Struct
    from acmil_b200.transmil import TransMIL
    torch.manual_seed(17)
    mod = TransMIL(Struct(D_feat=48, D_inner=d_inner, n_class=3)).to(dev()).eval()
    x = torch.randn(1, n, 48, generator=torch.Generator().manual_seed(n)).to(dev())
    with torch.no_grad():
        ref = mod(x)
    outs = _sharded_logits(mod, x, world)
    for y in outs:
        np.testing.assert_allclose(y.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_transmil_sharded_c3_full_size_against_the_reference():
    """BASELINE.json configs[2] as written: N = 50 000, dim 512, 256 landmarks, the bag sharded over 8 ranks -- against the
    logits of the reference itself (tests/golden/seeded_transmil_c3_n50000.npz)."""
    _, meta = load_golden("seeded_transmil_c3_n50000")
    mod = _seeded_transmil(meta).to(dev())
    x = golden_x(meta).to(dev())
    outs = _sharded_logits(mod, x, 8)
    for y in outs:
        np.testing.assert_allclose(y.cpu().numpy(), meta["out"], rtol=1e-3, atol=1e-4)
