"""CPU-side checks: the C-ABI library builds/loads and exports exactly what include/acmil_b200.h declares,
host-only entry points work without a GPU, compute entry points refuse loudly, and the drop-in modules keep
the reference's parameter names / shapes / initialisation order."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_names, load_golden

HEADERS = [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include"))) if f.endswith(".h")]


def declared_symbols():
    src = "\n".join(open(h).read() for h in HEADERS)
    return sorted(set(re.findall(r"ACMIL_API\s+[\w\s\*]+?\b(acmil_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import acmil_b200._lib as L
    lib = L.load()
    decl = declared_symbols()
    assert len(decl) >= 11
    assert sorted(L.SYMBOLS) == decl, "ctypes table and header disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (acmil_\w+)", nm))
    assert set(decl) <= exported
    assert lib.acmil_abi_version() == 1
    assert lib.acmil_device_count() >= 0


def test_struct_sizes_match_header_layout():
    import acmil_b200._lib as L
    assert C.sizeof(L.GpShape) == 48
    assert C.sizeof(L.GpWeights) == 64
    assert C.sizeof(L.GpBatch) == 64
    assert C.sizeof(L.GpHeads) == 48
    assert C.sizeof(L.GpOutputs) == 64
    assert C.sizeof(L.GpConsts) == (3 * 128 + 8 * 128 + 8 + 4) * 4 + 16
    # include/acmil_transmil.h
    assert C.sizeof(L.GemmDesc) == 7 * 8 + 4 * 4 + 10 * 8 + 2 * 4 + 8 + 3 * 4 + 2 * 4 + 3 * 4 + 5 * 8 + 8 + 2 * 4 + 2 * 8
    assert C.sizeof(L.NystromShape) == 16 * 4
    assert C.sizeof(L.NystromWeights) == 9 * 8
    assert C.sizeof(L.VitShape) == 16 * 4
    assert C.sizeof(L.VitBlockWeights) == 16 * 8
    assert C.sizeof(L.VitWeights) == 10 * 8


def test_transmil_host_entry_points_and_validation():
    import acmil_b200._lib as L
    lib = L.load()
    shape = L.NystromShape(1, 50001, 512, 8, 64, 256, 6, 1, 33, 0, 0, 1)
    n = C.c_size_t(0)
    assert lib.acmil_nystrom_workspace_bytes(C.byref(shape), C.byref(n)) == 0
    n_pad = 256 * 196
    assert n_pad >= 50001 and n.value >= (3 * 512 + 512 + 8 * 256) * n_pad * 4      # q, k, v^T, xn, one similarity
    bad = L.NystromShape(1, 100, 512, 8, 64, 254, 6, 1, 33, 0, 0, 1)
    assert lib.acmil_nystrom_workspace_bytes(C.byref(bad), C.byref(n)) == -1
    assert b"num_landmarks" in lib.acmil_last_error()
    # pre-split weight image: 256-byte header + fp16 hi / lo sections, rows padded to 8 halves, sections to 256 bytes
    assert lib.acmil_gemm_split_bytes(384, 1536, C.byref(n)) == 0 and n.value == 256 + 2 * 384 * 1536 * 2
    assert lib.acmil_gemm_split_bytes(3, 13, C.byref(n)) == 0 and n.value == 256 + 2 * 256
    assert lib.acmil_gemm_split_bytes(0, 13, C.byref(n)) == -1
    if not torch.cuda.is_available():      # no CPU path: compute entry points refuse loudly
        g = L.GemmDesc()
        g.a = g.b = g.c = 256
        g.m = g.n = g.k = 128
        g.batch = 1
        g.lda = g.ldb = g.ldc = 128
        assert lib.acmil_gemm_nt(C.byref(g), None) == -2
        assert lib.acmil_gemm_split_b(C.c_void_p(256), 8, 8, 8, C.c_void_p(256), 1 << 20, None) == -2
        assert lib.acmil_layernorm_rows(C.c_void_p(256), 8, 1, 8, None, None, 1e-5, C.c_void_p(256), 8, None) == -2


def test_host_only_entry_points_and_validation():
    import acmil_b200._lib as L
    lib = L.load()
    shape = L.GpShape(384, 128, 128, 5, 1, 0, 1, 0, 1, 1, 1, 0)
    n = C.c_size_t(0)
    assert lib.acmil_gp_packed_bytes(C.byref(shape), C.byref(n)) == 0
    f32 = (384 * 128 + 128 + 2 * 128 * 128 + 2 * 128 + 8 * 128 + 8) * 4
    assert n.value >= f32
    off = (C.c_int64 * 3)(0, 50000, 50007)
    batch = L.GpBatch(None, off, 2, 10, None, None, 50007)
    ws, part = C.c_size_t(0), C.c_size_t(0)
    assert lib.acmil_gp_sizes(C.byref(shape), C.byref(batch), L.IMPL_FFMA, C.byref(ws), C.byref(part)) == 0
    assert part.value == 2 * L.record_floats(5, 128, 10) * 4      # m, l, acc, cnt, score, idx, h (sections padded to 4)
    assert L.record_floats(5, 128, 10) == 8 + 8 + 640 + 8 + 52 + 52 + 6400
    assert ws.value > 0
    # validation errors come back as codes + messages, never exceptions/aborts
    bad = L.GpShape(384, 128, 128, 9, 1, 0, 1, 0, 1, 1, 1, 0)
    assert lib.acmil_gp_packed_bytes(C.byref(bad), C.byref(n)) == -1
    assert b"n_branch" in lib.acmil_last_error()
    off_bad = (C.c_int64 * 3)(0, 10, 5)
    batch_bad = L.GpBatch(None, off_bad, 2, 0, None, None, 10)
    assert lib.acmil_gp_sizes(C.byref(shape), C.byref(batch_bad), 0, C.byref(ws), C.byref(part)) == -1
    batch_bad = L.GpBatch(None, off, 2, 33, None, None, 50007)
    assert lib.acmil_gp_sizes(C.byref(shape), C.byref(batch_bad), 0, C.byref(ws), C.byref(part)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import acmil_b200._lib as L
    from acmil_b200 import ACMIL_GA, Struct
    lib = L.load()
    assert lib.acmil_device_count() == 0
    shape = L.GpShape(384, 128, 128, 5, 1, 0, 1, 0, 1, 1, 1, 0)
    w = L.GpWeights(1, None, 1, None, 1, None, 1, None)
    assert lib.acmil_gp_pack(C.byref(shape), C.byref(w), C.c_void_p(1), 1 << 30, None, None) == -2
    assert b"no CUDA device" in lib.acmil_last_error()
    m = ACMIL_GA(Struct(D_feat=384, D_inner=128, n_class=2, n_token=5), n_token=5).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 8, 384))


@pytest.mark.parametrize("name", golden_names("acmil_ga_") + golden_names("abmil_"))
def test_state_dict_layout_is_the_references(name):
    """Fixtures hold state_dicts saved from the reference's own modules: names, shapes and order must load."""
    from acmil_b200 import ABMIL, ACMIL_GA, Struct
    w, g = load_golden(name)
    d_feat, d_inner, n_class, k, n_masked = (int(v) for v in g["meta_conf"])
    conf = Struct(D_feat=d_feat, D_inner=d_inner, n_class=n_class, n_token=k)
    torch.manual_seed(int(g["meta_model_seed"]))
    m = ACMIL_GA(conf, n_token=k, n_masked_patch=n_masked) if name.startswith("acmil") else ABMIL(conf)
    sd = m.state_dict()
    assert list(sd.keys()) == list(w.keys())
    for key, v in sd.items():
        # same construction order => same initial values under the same seed as the reference
        assert np.array_equal(v.numpy(), w[key]), key


def test_attention_py_and_attmil_parameter_names():
    from acmil_b200.architecture import Attention as A
    from acmil_b200.architecture.attmil import AttentionGated, DAttention
    from test_oracle_golden import ATTMIL_AG_SHAPES, ATTMIL_DA_SHAPES
    w, _ = load_golden("attention_py_k4_n640")
    awc = A.Attention_with_Classifier(256, 128, 4, 3)
    assert list(awc.state_dict().keys()) == [k[len("awc::"):] for k in w if k.startswith("awc::")]
    for bias in (False, True):
        sd = AttentionGated(act="gelu", bias=bias).state_dict()
        assert {k: tuple(v.shape) for k, v in sd.items()} == ATTMIL_AG_SHAPES(bias)
        assert list(sd.keys()) == list(ATTMIL_AG_SHAPES(bias).keys())
    sd = DAttention(3, False, "relu").state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == ATTMIL_DA_SHAPES


def test_round2_entry_points_validate_and_refuse_without_a_gpu():
    """The entry points added in round 2 (backward row kernels, diversity loss, sharded Nystrom phases, log-sum-exp merge,
    PPEG rows): host-side validation gives codes + messages, and with no CUDA device every compute call returns
    ACMIL_E_CUDA -- there is no CPU path."""
    import acmil_b200._lib as L
    lib = L.load()
    gate_f, relu_f = C.c_int64(0), C.c_int64(0)
    assert lib.acmil_gp_bwd_workspace_floats(128, C.byref(gate_f), C.byref(relu_f)) == 0
    assert gate_f.value % (8 * 128 + 8 + 256) == 0 and relu_f.value % 128 == 0 and gate_f.value > 0
    shard = L.NystromShard(6304, 255, 512, 8, 64, 256, 32, 197, 6, 1, 33, 1, 0, 0, 1, 16)
    n = C.c_size_t(0)
    assert lib.acmil_nystrom_shard_workspace_bytes(C.byref(shard), C.byref(n)) == 0
    assert n.value >= (6304 * 512 + 2 * 8 * 6304 * 64 + 8 * 6304 * 256) * 4            # xn, q, k, one similarity
    bad = L.NystromShard(6300, 255, 512, 8, 64, 256, 32, 197, 6, 1, 33, 1, 0, 0, 1, 16)   # n_loc != m_loc * group_len
    assert lib.acmil_nystrom_shard_workspace_bytes(C.byref(bad), C.byref(n)) == -1
    assert b"n_loc" in lib.acmil_last_error()
    if torch.cuda.is_available():
        return
    p = C.c_void_p(256)
    args = L.GpBwdGateArgs(p, p, p, p, p, p, None, None, None, p, 8, 8, 0, 8, 128, 128, 5, 0, 1, 0, p, p, p, p, p)
    assert lib.acmil_gp_bwd_gate(C.byref(args), None) == -2
    assert lib.acmil_gp_bwd_relu_mask(p, p, 8, 128, None, p, 8, p, p, None) == -2
    assert lib.acmil_transpose_f32(p, 8, 8, 8, p, 8, None) == -2
    assert lib.acmil_div_loss_fwd(p, 8, 5, 8, p, p, p, None) == -2
    assert lib.acmil_div_loss_bwd(p, 8, 5, 8, p, p, p, p, 8, None) == -2
    assert lib.acmil_lse_merge(p, p, p, 2, 8, 64, 256, p, None) == -2
    assert lib.acmil_ppeg_fwd_rows(p, 4, 4, 32, p, p, p, p, p, p, p, 0, 4, None) == -2
    bufs = L.NystromShardBufs(p, None, p, p, p, p, p, p, p, p, p, p, p, p, 1 << 40)
    w = L.NystromWeights()
    assert lib.acmil_nystrom_shard_phase(C.byref(shard), C.byref(w), C.byref(bufs), 0, None) == -2
    assert b"no CUDA device" in lib.acmil_last_error()
