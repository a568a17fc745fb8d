"""TransMIL with the bag sharded over the GPUs of one box (BASELINE.json configs[2]): sequence-parallel
NystromAttention + PPEG with halo exchanges.  Same module, same weights, same result as ``TransMIL.forward``
(architecture/transMIL.py:60-91) -- the token sequence [cls, patches, wrap-around padding] is cut at the landmark-group
boundaries of nystrom_attention.py:95-114, so that every rank's landmark means are its own.

What crosses ranks, per Nystrom layer (csrc/tm_ops.cu, acmil_nystrom_shard_phase):
    landmark means q_l, k_l            all-gather   2 x [heads, m / P, d]      (1 MB in total at dim 512)
    pseudo-inverse of attn2            all-gather   [heads / P, m, m]          (heads are split over the ranks: 2 MB)
    attn3 v partial sums + (max, sum)  all-gather   [heads, m, d + 2]          (log-sum-exp merge = the softmax over ALL tokens)
    v halo of the depth-wise conv      neighbours   16 rows each way
and for PPEG: 3 grid rows + 3 tokens each way (transMIL.py:38-45, 7x7 depth-wise conv on the token grid).

``comm`` abstracts the exchanges: ``DistComm`` (torch.distributed: NCCL over NVLink on a box, gloo in CPU tests of the
host logic) or ``ThreadComm`` (the ranks are threads of one process on one GPU: how the GPU parity tests run the very
same per-rank code without a multi-GPU box).
"""
from __future__ import annotations

import ctypes as C
import math
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .transmil import gemm_nt, layernorm_rows


# ------------------------------------------------------------------------------------------ host logic (CPU-testable)
@dataclass(frozen=True)
class ShardPlan:
    """Who owns which token.  Tokens: 0 = cls, 1..n = patches, n+1..T-1 = the first patches again (transMIL.py:63-72);
    Nystrom pads the sequence at the FRONT to n_pad = m * l rows (nystrom_attention.py:72-80); rank p owns padded rows
    [p U, (p + 1) U), U = n_pad / P."""
    n: int
    world: int
    m: int

    def __post_init__(self):
        if self.m % self.world:
            raise ValueError(f"num_landmarks {self.m} must be divisible by the number of ranks {self.world}")
        if self.pad >= self.U:
            raise ValueError("sequence too short to shard: the front padding must fit inside rank 0's share")

    @property
    def g(self) -> int:               # side of the square token grid
        return int(np.ceil(np.sqrt(self.n)))

    @property
    def T(self) -> int:               # tokens incl. cls
        return 1 + self.g * self.g

    @property
    def l(self) -> int:
        return math.ceil(self.T / self.m)

    @property
    def n_pad(self) -> int:
        return self.T if self.T % self.m == 0 else self.m * self.l

    @property
    def pad(self) -> int:
        return self.n_pad - self.T

    @property
    def U(self) -> int:
        return self.n_pad // self.world

    @property
    def m_loc(self) -> int:
        return self.m // self.world

    def tokens(self, rank: int):
        """[t0, t1): the tokens of rank `rank`."""
        return max(rank * self.U - self.pad, 0), min((rank + 1) * self.U - self.pad, self.T)

    def lead_zero(self, rank: int) -> int:
        return self.pad if rank == 0 else 0

    def patch_rows(self, rank: int) -> np.ndarray:
        """Indices into the bag's n patch rows that feed this rank's tokens (cls excluded), in token order."""
        t0, t1 = self.tokens(rank)
        t = np.arange(max(t0, 1), t1)
        return np.where(t <= self.n, t - 1, t - self.n - 1)

    def ppeg_halo(self) -> int:
        return 3 * self.g + 3


class ThreadComm:
    """`world` ranks as threads of one process (single device): exchanges are hand-overs of tensor references."""

    def __init__(self, world: int):
        self.world = world
        self._bar = threading.Barrier(world)
        self._slots = [None] * world

    def all_gather(self, rank, t):
        self._slots[rank] = t
        self._bar.wait()
        out = list(self._slots)
        self._bar.wait()
        return out

    def shift(self, rank, to_left, to_right):
        """-> (what the left neighbour sent right, what the right neighbour sent left); None at the ends."""
        self._slots[rank] = (to_left, to_right)
        self._bar.wait()
        fl = self._slots[rank - 1][1] if rank > 0 else None
        fr = self._slots[rank + 1][0] if rank < self.world - 1 else None
        self._bar.wait()
        return fl, fr

    def broadcast0(self, rank, t):
        return self.all_gather(rank, t)[0]


class DistComm:
    """torch.distributed (NCCL on the GPUs of a box; gloo for CPU tests of the host logic)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group)

    def all_gather(self, rank, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        self._dist.all_gather(out, t.contiguous(), group=self.group)
        return out

    def shift(self, rank, to_left, to_right):
        dist = self._dist
        ops, fl, fr = [], None, None
        if rank > 0:
            fl = torch.empty_like(to_right)
            ops += [dist.P2POp(dist.isend, to_left.contiguous(), rank - 1, self.group), dist.P2POp(dist.irecv, fl, rank - 1, self.group)]
        if rank < self.world - 1:
            fr = torch.empty_like(to_left)
            ops += [dist.P2POp(dist.isend, to_right.contiguous(), rank + 1, self.group), dist.P2POp(dist.irecv, fr, rank + 1, self.group)]
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        return fl, fr

    def broadcast0(self, rank, t):
        self._dist.broadcast(t, src=0, group=self.group)
        return t


# ------------------------------------------------------------------------------------------ device side
def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def nystrom_shard_forward(attn, x_loc, plan: ShardPlan, rank: int, comm, *, ln=None, residual=None, n_out=None):
    """One rank's share of ``NystromAttention`` (+ pre-LayerNorm, + residual) over its tokens ``x_loc`` [t1 - t0, dim].
    n_out: None = all local rows, k > 0 = the first k local rows, 0 = this rank needs no output (it still takes part in
    the exchanges).  Returns [rows, dim] or None."""
    lib = L.load()
    dev = x_loc.device
    x_loc = x_loc.contiguous()
    heads, d, m, dim = attn.heads, attn.dim_head, attn.num_landmarks, x_loc.shape[1]
    world = comm.world
    if heads % world and world <= heads:
        raise ValueError(f"heads {heads} must be divisible by the number of ranks {world}")
    hc = heads // world if world <= heads else (1 if rank < heads else 0)
    hf = rank * hc if world <= heads else min(rank, heads)
    lead = plan.lead_zero(rank)
    n_loc = plan.U
    if x_loc.shape[0] != n_loc - lead:
        raise ValueError(f"rank {rank}: expected {n_loc - lead} local rows, got {x_loc.shape[0]}")
    halo = attn.conv_kernel // 2 if attn.residual else 0
    n_real = n_loc - lead
    rows = n_real if n_out is None else int(n_out)
    shard = L.NystromShard(n_loc, lead, dim, heads, d, m, plan.m_loc, plan.l, attn.pinv_iterations, int(attn.residual),
                           attn.conv_kernel if attn.residual else 1, attn._mode(), 0 if rows == n_real else max(rows, 1),
                           hf, hc, halo)
    nbytes = C.c_size_t(0)
    L.check(lib.acmil_nystrom_shard_workspace_bytes(C.byref(shard), C.byref(nbytes)))
    f32 = dict(device=dev, dtype=torch.float32)
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    inner = heads * d
    ql_loc, kl_loc = torch.empty(heads, plan.m_loc, d, **f32), torch.empty(heads, plan.m_loc, d, **f32)
    vt_ext = torch.zeros(inner, n_loc + 2 * halo, **f32)
    z = torch.empty(heads, m, m, **f32)
    kv_part, st_m, st_l = torch.empty(heads, d, m, **f32), torch.empty(heads * m, **f32), torch.empty(heads * m, **f32)
    kv = torch.empty(heads, d, m, **f32)
    out = torch.empty(max(rows, 1), dim, **f32)
    w = L.NystromWeights()
    keep = [attn.to_qkv.weight.contiguous(), attn.to_out[0].weight.contiguous(), attn.to_out[0].bias.contiguous()]
    if ln is not None:
        keep += [ln[0].contiguous(), ln[1].contiguous()]
        w.d_ln_w, w.d_ln_b, w.ln_eps = _ptr(keep[-2]), _ptr(keep[-1]), float(ln[2])
    w.d_wqkv, w.d_wout, w.d_bout = _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2])
    keep.extend(attn._split_ptrs(w))
    if attn.residual:
        keep.append(attn.res_conv.weight.contiguous())
        w.d_wconv = _ptr(keep[-1])
    res = None if residual is None else residual[:max(rows, 1)].contiguous()
    bufs = L.NystromShardBufs(_ptr(x_loc), _ptr(res), _ptr(out), _ptr(ql_loc), _ptr(kl_loc), None, None, _ptr(z), _ptr(kv_part),
                              _ptr(st_m), _ptr(st_l), _ptr(kv), _ptr(vt_ext), _ptr(ws), ws.numel())

    def phase(i):
        with torch.cuda.device(dev):
            L.check(lib.acmil_nystrom_shard_phase(C.byref(shard), C.byref(w), C.byref(bufs), i, _stream(dev)))

    phase(0)
    ql = torch.cat(comm.all_gather(rank, ql_loc), dim=1).contiguous()          # [heads, m, d]: rank order = landmark order
    kl = torch.cat(comm.all_gather(rank, kl_loc), dim=1).contiguous()
    bufs.d_ql, bufs.d_kl = _ptr(ql), _ptr(kl)
    if halo:
        fl, fr = comm.shift(rank, vt_ext[:, halo:2 * halo].contiguous(), vt_ext[:, n_loc:n_loc + halo].contiguous())
        if fl is not None:
            vt_ext[:, :halo] = fl
        if fr is not None:
            vt_ext[:, n_loc + halo:] = fr
    phase(1)
    if world > 1:
        zs = comm.all_gather(rank, z[hf:hf + hc].contiguous() if hc else z[:0].contiguous())
        z.copy_(torch.cat(zs, dim=0))
    phase(2)
    parts = torch.stack(comm.all_gather(rank, kv_part)).contiguous()
    sm = torch.stack(comm.all_gather(rank, st_m)).contiguous()
    sl = torch.stack(comm.all_gather(rank, st_l)).contiguous()
    with torch.cuda.device(dev):
        L.check(lib.acmil_lse_merge(_ptr(parts), _ptr(sm), _ptr(sl), world, heads, d, m, _ptr(kv), _stream(dev)))
    if rows == 0:
        return None
    phase(3)
    return out[:rows]


def ppeg_shard_forward(ppeg, h_loc, plan: ShardPlan, rank: int, comm):
    """PPEG (transMIL.py:38-45) on this rank's tokens: a full-size token buffer per rank in which only the own rows and the
    halo rows received from the neighbours are valid; the kernel runs on the grid rows the own tokens touch."""
    lib = L.load()
    dev = h_loc.device
    t0, t1 = plan.tokens(rank)
    g, T, dim = plan.g, plan.T, h_loc.shape[1]
    hl = min(plan.ppeg_halo(), h_loc.shape[0])
    buf = torch.zeros(T, dim, device=dev, dtype=torch.float32)
    buf[t0:t1] = h_loc
    fl, fr = comm.shift(rank, h_loc[:hl].contiguous(), h_loc[-hl:].contiguous())
    if fl is not None:
        buf[t0 - fl.shape[0]:t0] = fl
    if fr is not None:
        buf[t1:t1 + fr.shape[0]] = fr
    out = torch.empty_like(buf)
    first = max(t0, 1) - 1                      # grid positions of the own tokens: [first, last]
    last = t1 - 2
    if last >= first:
        y0, y1 = first // g, last // g + 1
        ps = [t.contiguous() for t in (ppeg.proj.weight, ppeg.proj.bias, ppeg.proj1.weight, ppeg.proj1.bias, ppeg.proj2.weight,
                                       ppeg.proj2.bias)]
        with torch.cuda.device(dev):
            L.check(lib.acmil_ppeg_fwd_rows(_ptr(buf), g, g, dim, *[_ptr(t) for t in ps], _ptr(out), y0, y1 - y0, _stream(dev)))
    if t0 == 0:
        out[0] = buf[0]                         # the class token passes through
    return out[t0:t1].contiguous()


@torch.no_grad()
def transmil_forward_sharded(model, x_rows, n: int, rank: int, comm):
    """``TransMIL.forward`` for one bag of n patches whose rows are spread over ``comm.world`` ranks.
    x_rows [len(plan.patch_rows(rank)), D_feat]: the patch rows of this rank's tokens (ShardPlan.patch_rows).
    Returns the logits [1, n_class] on every rank."""
    plan = ShardPlan(n, comm.world, model.layer1.attn.num_landmarks)
    dev = x_rows.device
    fc1 = model._fc1[0]
    t0, t1 = plan.tokens(rank)
    D = fc1.out_features
    h = torch.empty(t1 - t0, D, device=dev, dtype=torch.float32)
    first = 1 if t0 == 0 else 0
    if x_rows.shape[0] != t1 - t0 - first:
        raise ValueError(f"rank {rank}: expected {t1 - t0 - first} patch rows, got {x_rows.shape[0]}")
    if x_rows.shape[0]:
        gemm_nt(x_rows.contiguous(), fc1.weight, bias=fc1.bias, relu=True, out=h[first:],        # _fc1 (:61)
                b_split=model._split.get("fc1", fc1.weight))
    if t0 == 0:
        h[0] = model.cls_token[0, 0]
    l1, l2 = model.layer1, model.layer2
    h = nystrom_shard_forward(l1.attn, h, plan, rank, comm, ln=(l1.norm.weight, l1.norm.bias, l1.norm.eps), residual=h)      # (:75)
    h = ppeg_shard_forward(model.pos_layer, h, plan, rank, comm)                                                             # (:78)
    cls = nystrom_shard_forward(l2.attn, h, plan, rank, comm, ln=(l2.norm.weight, l2.norm.bias, l2.norm.eps), residual=h,
                                n_out=1 if rank == 0 else 0)                                                                  # (:81, 84)
    logits = torch.empty(1, model._fc2.out_features, device=dev, dtype=torch.float32)
    if rank == 0:
        c = layernorm_rows(cls[:1], model.norm.weight, model.norm.bias, model.norm.eps)
        logits = gemm_nt(c, model._fc2.weight, bias=model._fc2.bias)                                                          # (:87)
    return comm.broadcast0(rank, logits)


def run_threads(world: int, fn):
    """Runs fn(rank, comm) for rank = 0..world-1 as threads sharing one ThreadComm; returns the per-rank results."""
    comm = ThreadComm(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = fn(r, comm)
        except BaseException as e:      # noqa: BLE001  (re-raised below; a dead thread must not leave the others at a barrier)
            err[r] = e
            comm._bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out
