"""CPU restatement of the arithmetic behind two GEMM-engine features (acmil_b200/csrc/tm_gemm.cu), so that what the kernels
compute is pinned without a GPU: the fp16 hi / lo split of the weight products (acmil_gemm_split_b + split_h2) and the
chunked softmax that rides on two products (acmil_gemm_desc.softmax_stats_out / _in).  The GPU tests compare the kernels with
fp64 / torch.softmax directly (tests/test_transmil_gpu.py); these show that the formulation itself is fp32-faithful."""
import numpy as np
import pytest


def split_h2(x):
    """fp32 -> (hi, lo) as the kernels do: hi = the top 11 significant bits (13 low mantissa bits cleared; exact in fp16 for
    normal values), lo = the rest rounded to fp16; both saturate at the fp16 maximum."""
    x = np.asarray(x, np.float32)
    hi32 = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    with np.errstate(over="ignore"):
        hi = np.clip(hi32, -65504, 65504).astype(np.float16)
        lo = np.clip(x - hi32, -65504, 65504).astype(np.float16)
    return hi, lo


def weight_scale(w):
    """the power of two acmil_gemm_split_b picks on the device: max |w| * scale in [2^13, 2^14)"""
    mx = float(np.abs(w).max())
    if not (mx > 0 and np.isfinite(mx)):
        return 1.0
    _, e = np.frexp(np.float32(mx))
    return float(np.ldexp(1.0, int(np.clip(14 - e, -100, 100))))


@pytest.mark.parametrize("wscale", [1e-4, 0.02, 1.0, 50.0])
def test_fp16_split_products_are_fp32_faithful(wscale):
    rng = np.random.default_rng(3)
    m, n, k = 96, 80, 384
    a = rng.standard_normal((m, k)).astype(np.float32)
    a[::7] *= 30
    a[1::5] *= 1e-3
    w = (rng.standard_normal((n, k)) * wscale).astype(np.float32)
    s = weight_scale(w)
    assert 2 ** 13 <= np.abs(w).max() * s < 2 ** 14
    a_hi, a_lo = split_h2(a)
    b_hi, b_lo = split_h2(w * np.float32(s))                       # the scaling is a power of two: exact
    f = np.float64
    got = (a_lo.astype(f) @ b_hi.astype(f).T + a_hi.astype(f) @ b_lo.astype(f).T + a_hi.astype(f) @ b_hi.astype(f).T) / s
    ref = a.astype(f) @ w.astype(f).T
    bound = (np.abs(a).astype(f) + 0.125) @ np.abs(w).astype(f).T  # the lo part of an |a| < 0.25 keeps an absolute 2^-25
    err = float((np.abs(got - ref) / bound).max())
    assert err < 2 ** -21, err                                     # dropped a_lo b_lo term + fp16 rounding of the lo parts
    # hi parts of normal fp16 magnitudes are exactly representable: nothing is lost before the lo parts
    normal = np.abs(a) >= 2.0 ** -14
    assert np.array_equal(a_hi.astype(np.float32)[normal], (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)[normal])
    # one plain fp16 product (no lo parts) is three orders of magnitude coarser: the split is what buys the accuracy
    coarse = a_hi.astype(f) @ b_hi.astype(f).T / s
    assert float((np.abs(coarse - ref) / bound).max()) > 100 * err


def test_fp16_split_saturates_instead_of_overflowing():
    hi, lo = split_h2(np.array([1e6, -3e5, 65504.0, 7e4], np.float32))
    assert np.all(np.isfinite(hi.astype(np.float32))) and float(np.abs(hi.astype(np.float32)).max()) == 65504.0
    assert weight_scale(np.zeros((2, 2), np.float32)) == 1.0


@pytest.mark.parametrize("n", [197, 256, 40, 1000])
def test_chunked_softmax_identity(n):
    """exp(s - max_c) * exp(max_c - M) / sum_c' l_c' exp(max_c' - M) == softmax(s), chunk by chunk of 32 columns"""
    rng = np.random.default_rng(n)
    s = (rng.standard_normal((50, n)) * 8).astype(np.float32)
    s[::3] *= 4                                                    # rows whose chunks differ by e^100 in magnitude
    nch = (n + 31) // 32
    pad = np.full((s.shape[0], nch * 32), -np.inf, np.float32)
    pad[:, :n] = s
    ch = pad.reshape(s.shape[0], nch, 32)
    mc = ch.max(-1)                                                # what the score product's epilogue stores per chunk ...
    e = np.exp(ch - mc[..., None], dtype=np.float32)               # ... next to these numerators (0 in the padding)
    lc = e.sum(-1, dtype=np.float32)
    M = mc.max(-1, keepdims=True)                                  # what the next product's converter derives per row
    L = (lc * np.exp(mc - M, dtype=np.float32)).sum(-1, keepdims=True, dtype=np.float32)
    f = np.exp(mc - M, dtype=np.float32) / L
    got = (e * f[..., None]).reshape(s.shape[0], -1)[:, :n]
    s64 = s.astype(np.float64)
    ref = np.exp(s64 - s64.max(-1, keepdims=True))
    ref /= ref.sum(-1, keepdims=True)
    assert float(np.abs(got - ref).max()) < 5e-7
    assert float(np.abs(got.sum(-1) - 1).max()) < 1e-5
