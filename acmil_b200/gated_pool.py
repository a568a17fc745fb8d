"""Host-side driver of the fused gated-attention pool (the C-ABI calls of include/acmil_b200.h).

PyTorch is used here for device memory, streams and (when a bag is sharded over ranks)
``torch.distributed``; all arithmetic of the path runs in libacmil_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _lib as L


@dataclass(frozen=True)
class GatedPoolSpec:
    d_in: int
    d_inner: int
    n_branch: int = 1
    d_attn: int = 128
    front: bool = True          # h = act(x W1^T [+ b1]) before the gate (DimReduction / attmil feature)
    front_bias: bool = False
    front_act: str = "relu"
    act_a: str = "tanh"
    gated: bool = True
    gate_bias: bool = True
    score_bias: bool = True

    def c_struct(self) -> L.GpShape:
        return L.GpShape(self.d_in, self.d_inner, self.d_attn, self.n_branch, int(self.front),
                         int(self.front_bias), L.ACT_IDS[self.front_act], L.ACT_IDS[self.act_a],
                         int(self.gated), int(self.gate_bias), int(self.score_bias), 0)


@dataclass
class GatedPoolResult:
    sub: Optional[torch.Tensor]        # [S, K, C]
    slide: Optional[torch.Tensor]      # [S, C]
    afeat: torch.Tensor                # [S, K, L]
    bag_feat: torch.Tensor             # [S, L]
    lse_m: torch.Tensor                # [S, K]
    lse_l: torch.Tensor                # [S, K]
    scores: Optional[torch.Tensor]     # [K, R_local] raw scores, -1e9 at masked positions
    topk_idx: Optional[torch.Tensor]   # [S, K, n_masked] int64
    masked_idx: Optional[torch.Tensor]  # [S, K, keep_ld] int64
    row_offsets: list = field(default_factory=list)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"acmil_b200: {name} must be a CUDA tensor (there is no CPU path)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class GatedPool:
    """One gated-attention pooling head: packs the weights once, then runs bags through the kernels."""

    def __init__(self, spec: GatedPoolSpec, impl: int = L.IMPL_AUTO):
        if not (1 <= spec.n_branch <= L.MAX_BRANCH):
            raise ValueError(f"n_branch must be in [1, {L.MAX_BRANCH}]")
        self.spec = spec
        self.impl = impl
        self._shape = spec.c_struct()
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._w1 = None           # (w1, b1, (wv, bv, wu, bu, ww, bw)) fp32 copies of the last pack()
        self._tail: Optional["GatedPool"] = None
        self._wcat = self._wcat_key = None
        self._imgs: dict = {}     # fp16 hi / lo images of w1 and [wv; wu] for the fp16-split GEMM kernel
        self._bufs: dict = {}

    # ------------------------------------------------------------------ weights
    def pack(self, w1, b1, wv, bv, wu, bu, ww, bw) -> torch.Tensor:
        """Packs nn.Linear-layout weights ([out, in]) into the kernels' layouts.  Re-packs only when
        a tensor's storage or in-place version changed (optimizer steps bump ``_version``)."""
        tensors = (w1, b1, wv, bv, wu, bu, ww, bw)
        key = tuple((None if t is None else (t.data_ptr(), t._version, tuple(t.shape))) for t in tensors)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        lib = L.load()
        dev = wv.device
        _require_cuda(wv, "weights")
        nbytes = C.c_size_t(0)
        L.check(lib.acmil_gp_packed_bytes(C.byref(self._shape), C.byref(nbytes)))
        keep = [None if t is None else _f32c(t) for t in tensors]
        w = L.GpWeights(*[_ptr(t) for t in keep])
        packed = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            # consts = NULL: the kernels' constants are written into the blob on the device; no host copy, no stream sync
            L.check(lib.acmil_gp_pack(C.byref(self._shape), C.byref(w), _ptr(packed), nbytes.value, None, C.c_void_p(st)))
        self._packed, self._packed_key = packed, key
        self._keepalive = keep
        self._w1 = None if keep[0] is None else (keep[0], keep[1], tuple(keep[2:]))      # for _split_front
        self._wcat = self._wcat_key = None
        self._imgs = {}
        return packed

    # ------------------------------------------------------------------ front projection on the GEMM engine
    def _split_front(self, x, n_masked, impl):
        """-> (tail GatedPool, its packed weights, h) when this call should run as  h = act(x W1^T + b1)  on the tcgen05 GEMM
        engine (3xTF32, fp32-faithful) followed by the pool kernels on h; None = run fused as is.  Only without masking: the
        mask indices are promised bit-exact against the fp32 reference, which the exact FFMA kernel guarantees and a
        2^-21-accurate GEMM in front of it would not in near-ties; eval-mode outputs have the 1e-3 bar."""
        sp = self.spec
        impl = self.impl if impl is None else impl
        if (not sp.front or n_masked > 0 or impl != L.IMPL_AUTO or self._w1 is None or x.shape[0] < 4096
                or L.load().acmil_gp_umma_supported(C.byref(self._shape))):
            return None
        from .transmil import SplitImage, gemm_mode, gemm_nt
        if self._tail is None:
            tail_spec = GatedPoolSpec(d_in=sp.d_inner, d_inner=sp.d_inner, n_branch=sp.n_branch, d_attn=sp.d_attn, front=False,
                                      act_a=sp.act_a, gated=sp.gated, gate_bias=sp.gate_bias, score_bias=sp.score_bias)
            self._tail = GatedPool(tail_spec, L.IMPL_FFMA)
        w1, b1, rest = self._w1
        packed2 = self._tail.pack(None, None, *rest)
        xf = x if x.dtype == torch.float32 else x.float()
        def image(name, w):      # the weights are the B operands: pre-split once per pack()
            if gemm_mode() != 2:
                return None
            if name not in self._imgs:
                self._imgs[name] = SplitImage(w)
            return self._imgs[name]

        h = gemm_nt(xf, w1, bias=b1, relu=sp.front_act == "relu", gelu=sp.front_act == "gelu", b_split=image("w1", w1))
        # ... and the gate products h Wv^T | h Wu^T too (the kernel adds the biases, applies the gate and pools)
        wv, _bv, wu, _bu = rest[0], rest[1], rest[2], rest[3]
        if self._wcat is None or self._wcat_key != (wv.data_ptr(), None if wu is None else wu.data_ptr()):
            self._wcat = (torch.cat([wv, wu], 0) if sp.gated else wv).contiguous()
            self._wcat_key = (wv.data_ptr(), None if wu is None else wu.data_ptr())
            self._imgs.pop("wcat", None)
        zz = gemm_nt(h, self._wcat, b_split=image("wcat", self._wcat))
        return self._tail, packed2, h, zz

    def invalidate(self) -> None:
        """Forget the packed weights.  The cache key is (storage pointer, in-place version, shape) of every weight tensor:
        optimizer steps, ``load_state_dict`` and ``copy_`` bump the version, but writes through ``param.data`` (EMA / manual
        weight surgery, as in the reference's attmil initialiser) do not -- call this after such a write."""
        self._packed = self._packed_key = None

    def _buf(self, name: str, nbytes: int, dev) -> torch.Tensor:
        b = self._bufs.get((name, dev))
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
            self._bufs[(name, dev)] = b
        return b

    # ------------------------------------------------------------------ forward
    def partial(self, packed: torch.Tensor, x: torch.Tensor, row_offsets: Sequence[int], *, n_masked: int = 0,
                want_scores: bool = True, shard_begin: Optional[Sequence[int]] = None, impl: Optional[int] = None,
                exchange=None, z: Optional[torch.Tensor] = None):
        """Row pass over this device's rows.  Returns (record, ctx): ``record`` is the flat fp32 partial
        record (what sharded ranks all-gather), ``ctx`` carries the batch description for finish().
        With ``exchange`` (a sharding.PeerExchange) the reduce kernel stores the records into every peer's gather buffer
        over NVLink instead and ``record`` is None."""
        lib = L.load()
        sp = self.spec
        _require_cuda(x, "x")
        if x.dtype not in (torch.float32, torch.float16) or not x.is_contiguous() or x.dim() != 2 or x.shape[1] != sp.d_in:
            raise ValueError(f"x must be a contiguous fp32 (or fp16: the H5 storage dtype, widened exactly inside the kernels) "
                             f"[R, {sp.d_in}] tensor")
        dev = x.device
        S = len(row_offsets) - 1
        R = int(row_offsets[-1])
        if R != x.shape[0]:
            raise ValueError("row_offsets[-1] must equal the number of rows of x")
        impl = self.impl if impl is None else impl
        off = (C.c_int64 * (S + 1))(*[int(v) for v in row_offsets])
        sb = (C.c_int64 * max(S, 1))(*[int(v) for v in shard_begin]) if shard_begin is not None else None
        scores = torch.empty((sp.n_branch, max(R, 1)), dtype=torch.float32, device=dev) if want_scores else None
        batch = L.GpBatch(_ptr(x), off, S, int(n_masked), sb, _ptr(scores), max(R, 1), int(x.dtype == torch.float16), 0, _ptr(z))
        ws_b, part_b = C.c_size_t(0), C.c_size_t(0)
        L.check(lib.acmil_gp_sizes(C.byref(self._shape), C.byref(batch), impl, C.byref(ws_b), C.byref(part_b)))
        ws = self._buf("ws", ws_b.value, dev)
        part = None if exchange is not None else torch.empty(max(part_b.value, 4) // 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            consts = None      # (acmil_gp_consts is an optional host copy; the kernels read the device copy in the blob)
            if exchange is not None:
                xs = exchange.c_struct(part_b.value)
                L.check(lib.acmil_gp_partial_x(C.byref(self._shape), _ptr(packed), consts, C.byref(batch), impl, _ptr(ws),
                                               ws.numel(), C.byref(xs), st))
            else:
                L.check(lib.acmil_gp_partial(C.byref(self._shape), _ptr(packed), consts, C.byref(batch), impl, _ptr(ws),
                                             ws.numel(), _ptr(part), part.numel() * 4, st))
        ctx = dict(batch=batch, keepalive=(off, sb, x, z), scores=scores, S=S, R=R, dev=dev, n_masked=int(n_masked),
                   row_offsets=list(row_offsets), ws=ws, impl=impl,
                   exchange=exchange, partial_bytes=part_b.value)
        return part, ctx

    def rescued_bags(self, ctx: dict) -> list:
        """Diagnostics: per bag of the batch behind ``ctx``, 1 when the tcgen05 kernel's bounded candidate scratch ran out
        (e.g. rows sorted by ascending score) and the exact FFMA kernel redid the bag on the device.  Synchronises."""
        lib = L.load()
        S = ctx["S"]
        flags = (C.c_int32 * max(S, 1))()
        with torch.cuda.device(ctx["dev"]):
            st = C.c_void_p(torch.cuda.current_stream(ctx["dev"]).cuda_stream)
            L.check(lib.acmil_gp_overflow_flags(C.byref(self._shape), C.byref(ctx["batch"]), ctx["impl"], _ptr(ctx["ws"]),
                                                flags, st))
        return [int(flags[i]) for i in range(S)]

    def finish(self, ctx: dict, records: torch.Tensor, n_ranks: int = 1, *, keep: Optional[Sequence[int]] = None,
               rsel: Optional[torch.Tensor] = None, branch_w=None, branch_b=None, head_w=None, head_b=None,
               slide_head: bool = False, shared_head: bool = False, rand: Optional[torch.Tensor] = None) -> GatedPoolResult:
        """Merges ``n_ranks`` partial records (back to back in ``records``) and produces the outputs.

        The mask draw comes either sorted (``rsel`` = ``argsort(rand)[..., :keep]``, int64) or raw (``rand`` =
        ``torch.rand(..., n)`` itself, [S, K, n] fp32: the kernel sorts, no extra launches)."""
        lib = L.load()
        sp = self.spec
        K, Lw = sp.n_branch, sp.d_inner
        S, R, dev, n_masked = ctx["S"], ctx["R"], ctx["dev"], ctx["n_masked"]
        C_ = branch_w.shape[1] if branch_w is not None else (head_w.shape[0] if head_w is not None else 0)
        keep = [int(v) for v in keep] if keep is not None else [0] * S
        keep_ld = max([1] + keep)
        keep_arr = (C.c_int32 * max(S, 1))(*keep)
        rand_ld = 0
        if n_masked > 0 and any(keep):
            if rand is not None:
                rand = rand.to(device=dev, dtype=torch.float32).contiguous()
                rand_ld = int(rand.shape[-1])
                if rand.numel() != S * K * rand_ld:
                    raise ValueError("rand must be [S, K, n] uniform draws")
                rsel = None
            elif rsel is None:
                raise ValueError("masking needs rsel or rand")
            else:
                rsel = rsel.to(device=dev, dtype=torch.int64).contiguous()
                if rsel.numel() < S * K * keep_ld:
                    raise ValueError("rsel must hold [S, K, max(keep)] indices")
        else:
            rsel = rand = None
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        out_sub = f(S, K, C_) if (branch_w is not None or shared_head) else None
        out_slide = f(S, C_) if slide_head else None
        afeat, bag, lse_m, lse_l = f(S, K, Lw), f(S, Lw), f(S, K), f(S, K)
        topk = torch.empty((S, K, n_masked), dtype=torch.int64, device=dev) if n_masked > 0 else None
        masked = torch.full((S, K, keep_ld), -1, dtype=torch.int64, device=dev) if n_masked > 0 else None
        bw_ = None if branch_w is None else _f32c(branch_w)
        bb_ = None if branch_b is None else _f32c(branch_b)
        hw_ = None if head_w is None else _f32c(head_w)
        hb_ = None if head_b is None else _f32c(head_b)
        heads = L.GpHeads(C_, K if branch_w is not None else 0, _ptr(bw_), _ptr(bb_), int(slide_head),
                          int(shared_head), _ptr(hw_), _ptr(hb_))
        outs = L.GpOutputs(_ptr(out_sub), _ptr(out_slide), _ptr(afeat), _ptr(bag), _ptr(lse_m), _ptr(lse_l),
                           _ptr(topk), _ptr(masked))
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if ctx.get("exchange") is not None:
                xs = ctx["exchange"].c_struct(ctx["partial_bytes"])
                L.check(lib.acmil_gp_finish_x(C.byref(self._shape), C.byref(ctx["batch"]), C.byref(xs), keep_arr, _ptr(rsel),
                                              _ptr(rand), rand_ld, keep_ld, C.byref(heads), C.byref(outs), st))
            elif rand is not None:
                L.check(lib.acmil_gp_finish_rand(C.byref(self._shape), C.byref(ctx["batch"]), _ptr(records), records.numel() * 4,
                                                 int(n_ranks), keep_arr, _ptr(rand), rand_ld, keep_ld, C.byref(heads),
                                                 C.byref(outs), st))
            else:
                L.check(lib.acmil_gp_finish(C.byref(self._shape), C.byref(ctx["batch"]), _ptr(records), records.numel() * 4,
                                            int(n_ranks), keep_arr, _ptr(rsel), keep_ld, C.byref(heads), C.byref(outs), st))
        scores = ctx["scores"]
        return GatedPoolResult(out_sub, out_slide, afeat, bag, lse_m, lse_l,
                               None if scores is None else scores[:, :R], topk, masked, ctx["row_offsets"])

    def run(self, packed: torch.Tensor, x: torch.Tensor, row_offsets: Sequence[int], *, n_masked: int = 0,
            keep: Optional[Sequence[int]] = None, rsel: Optional[torch.Tensor] = None,
            branch_w: Optional[torch.Tensor] = None, branch_b: Optional[torch.Tensor] = None,
            head_w: Optional[torch.Tensor] = None, head_b: Optional[torch.Tensor] = None,
            slide_head: bool = False, shared_head: bool = False, want_scores: bool = True,
            shard_begin: Optional[Sequence[int]] = None, group=None, impl: Optional[int] = None,
            rand: Optional[torch.Tensor] = None, exchange=None, z: Optional[torch.Tensor] = None) -> GatedPoolResult:
        """x: [R, d_in] fp32 CUDA, rows of S bags concatenated; row_offsets: S+1 host ints.

        With ``group`` (a torch.distributed process group) every rank passes its row shard of each bag and
        ``shard_begin`` (global index of its first row per bag); the per-bag partial records (a few KB)
        are all-gathered over NCCL and every rank finishes redundantly -- no other exchange.
        """
        split = self._split_front(x, n_masked, impl) if z is None else None
        if split is not None:
            # shapes outside the fused tcgen05 kernel (D_inner 256 / 384 / 512, front-layer bias, GELU front): the front
            # projection -- most of the FLOPs -- runs on the tcgen05 GEMM engine and the exact FFMA kernel does the rest
            op2, packed2, h, zz = split
            return op2.run(packed2, h, row_offsets, n_masked=0, keep=keep, branch_w=branch_w, branch_b=branch_b, head_w=head_w,
                           head_b=head_b, slide_head=slide_head, shared_head=shared_head, want_scores=want_scores,
                           shard_begin=shard_begin, group=group, exchange=exchange, z=zz)
        if exchange is not None:
            # records travel inside the kernels (NVLink stores + flags): no collective, graph-capturable
            _, ctx = self.partial(packed, x, row_offsets, n_masked=n_masked, want_scores=want_scores,
                                  shard_begin=shard_begin, impl=impl, exchange=exchange, z=z)
            return self.finish(ctx, None, exchange.world, keep=keep, rsel=rsel, branch_w=branch_w, branch_b=branch_b,
                               head_w=head_w, head_b=head_b, slide_head=slide_head, shared_head=shared_head, rand=rand)
        part, ctx = self.partial(packed, x, row_offsets, n_masked=n_masked, want_scores=want_scores,
                                 shard_begin=shard_begin, impl=impl, z=z)
        from .sharding import gather_records
        world = torch.distributed.get_world_size(group) if (group is not None or (
            shard_begin is not None and torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        gathered = gather_records(part, group) if world > 1 else part
        return self.finish(ctx, gathered, world, keep=keep, rsel=rsel, branch_w=branch_w, branch_b=branch_b,
                           head_w=head_w, head_b=head_b, slide_head=slide_head, shared_head=shared_head, rand=rand)

    # ------------------------------------------------------------------ small ops
    @staticmethod
    def attn_stats(scores: torch.Tensor, row_offsets: Sequence[int], lse_m: torch.Tensor, lse_l: torch.Tensor):
        """(gram [S,K,K], ent [S,K], div [S]) of softmax(scores) per bag
        (Step3_WSI_classification_ACMIL.py:208-214 and :259)."""
        lib = L.load()
        _require_cuda(scores, "scores")
        K = scores.shape[0]
        S = len(row_offsets) - 1
        dev = scores.device
        if scores.stride(1) != 1:
            scores = scores.contiguous()
        off = (C.c_int64 * (S + 1))(*[int(v) for v in row_offsets])
        gram = torch.empty((S, K, K), dtype=torch.float32, device=dev)
        ent = torch.empty((S, K), dtype=torch.float32, device=dev)
        div = torch.empty((S,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            L.check(lib.acmil_gp_attn_stats(_ptr(scores), scores.stride(0), K, off, S, _ptr(lse_m.contiguous()),
                                            _ptr(lse_l.contiguous()), _ptr(gram), _ptr(ent), _ptr(div), st))
        return gram, ent, div

    @staticmethod
    def softmax_rows(a: torch.Tensor) -> torch.Tensor:
        """F.softmax(a, dim=1) for a [K, N] fp32 CUDA matrix (Attention.py:56-57)."""
        lib = L.load()
        _require_cuda(a, "a")
        if a.stride(1) != 1:
            a = a.contiguous()
        out = torch.empty((a.shape[0], a.shape[1]), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            st = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
            L.check(lib.acmil_softmax_rows(_ptr(a), a.stride(0), a.shape[0], a.shape[1], _ptr(out), out.stride(0), st))
        return out
