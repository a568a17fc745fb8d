"""Bag sharding over the GPUs of one box: each rank owns a contiguous run of rows of every bag.

The only exchange on the path is an all-gather of the per-bag partial records (K*(L+2) floats, plus
the top-n candidate rows in training mode): a few KB per bag over NCCL/NVLink.  Every rank then
finishes redundantly, so there is no second collective and no designated root.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int) -> List[int]:
    """Row r of the bag lives on the rank with bounds[rank] <= r < bounds[rank + 1]."""
    return [n_rows * r // world for r in range(world + 1)]


def gather_records(part: torch.Tensor, group=None) -> torch.Tensor:
    """All-gathers one flat record tensor per rank into [world * numel] (rank-major)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return part
    out = torch.empty(world * part.numel(), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(out, part.contiguous(), group=group)
    return out


class PeerExchange:
    """Peer-memory exchange of the per-bag partial records (include/acmil_b200.h: acmil_gp_exchange).

    Every rank owns one symmetric-memory buffer ``[flags | 2 parities x world x records]`` that all peers have mapped
    (torch.distributed._symmetric_memory: CUDA VMM handles exchanged over the process group, NVLink peer access).  The
    reduce kernel of a rank stores its records into the buffers of ALL ranks and raises its flag there; the finish kernel
    waits for the flags in-kernel.  PyTorch only allocates and maps the memory; no collective is launched per step.
    """
    FLAG_BYTES = 256

    def __init__(self, max_partial_bytes: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise ValueError("PeerExchange supports up to 16 ranks")
        self.capacity = int(max_partial_bytes)
        self.gather_bytes = 2 * self.world * self.capacity
        n = (self.FLAG_BYTES + self.gather_bytes + 3) // 4
        self.buf = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.state = torch.zeros(8, dtype=torch.int32, device=device)      # [0] epoch, [4..7] tickets
        torch.cuda.synchronize(device)
        dist.barrier(group)      # every rank's flags are zero before anybody can raise one

    def c_struct(self, partial_bytes: int):
        """acmil_gp_exchange for a step whose per-rank records take ``partial_bytes``."""
        import ctypes as C
        from . import _lib as L
        if partial_bytes > self.capacity:
            raise ValueError(f"PeerExchange sized for {self.capacity} bytes per rank, step needs {partial_bytes}")
        x = L.GpExchange()
        x.n_ranks, x.rank = self.world, self.rank
        for r, p in enumerate(self.ptrs):
            x.d_flags[r] = p
            x.d_gather[r] = p + self.FLAG_BYTES
        x.d_epoch = self.state.data_ptr()
        x.d_ticket = self.state.data_ptr() + 16
        x.gather_bytes = self.gather_bytes
        return x


def draw_rsel(k: int, nm: int, keep: int, device, group=None, src: int = 0) -> torch.Tensor:
    """The reference's draw (transformer.py:316) made on rank `src` and broadcast, so that every rank
    masks the same patches: argsort(rand(k, nm))[:, :keep]."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if rank == src or world == 1:
        rsel = torch.argsort(torch.rand(k, nm, device=device), dim=-1)[:, :keep].contiguous()
    else:
        rsel = torch.empty((k, keep), dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(rsel, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return rsel


class ShardedACMIL:
    """Runs an acmil_b200.ACMIL_GA on a bag whose rows are spread over the ranks of `group`.

    forward(x_local [1, n_local, D_feat], n_total, row_begin) -> (sub [K, C], slide [1, C],
    A_local [1, K, n_local]) -- logits identical on every rank, scores stay sharded.
    """

    def __init__(self, model, group=None, exchange: Optional["PeerExchange"] = None):
        self.model = model
        self.group = group
        self.exchange = exchange      # records over peer memory inside the kernels instead of an NCCL all-gather

    @torch.no_grad()
    def forward_bags(self, x_local_cat: torch.Tensor, local_offsets, n_total, row_begin):
        """S sharded bags in one launch: this rank's rows of every bag back to back, ``n_total[s]`` rows per whole bag,
        ``row_begin[s]`` the global index of the rank's first row.  -> (sub [S, K, C], slide [S, C], local scores)."""
        return self.model.forward_bags(x_local_cat, local_offsets, shard_begin=list(row_begin), n_total=list(n_total),
                                       exchange=self.exchange, group=None if self.exchange is not None else self.group)

    @torch.no_grad()
    def forward(self, x_local: torch.Tensor, n_total: int, row_begin: int, use_mask: Optional[bool] = None):
        m = self.model
        op = m._op
        use_mask = m.training if use_mask is None else use_mask
        k = m.attention.K
        n_masked = keep = 0
        rsel = None
        if m.n_masked_patch > 0 and use_mask:
            nm = min(m.n_masked_patch, n_total)
            keep = int(nm * m.mask_drop)
            rsel = draw_rsel(k, nm, keep, x_local.device, self.group)
            n_masked = m.n_masked_patch if keep > 0 else 0
        w = m._weights()
        packed = op.pack(w.get("w1"), w.get("b1"), w["wv"], w.get("bv"), w.get("wu"), w.get("bu"), w["ww"], w.get("bw"))
        x2d = x_local[0].to(torch.float32).contiguous()
        res = op.run(packed, x2d, [0, x2d.shape[0]], n_masked=n_masked, keep=[keep], rsel=rsel,
                     branch_w=torch.stack([c.fc.weight for c in m.classifier]),
                     branch_b=torch.stack([c.fc.bias for c in m.classifier]),
                     head_w=m.Slide_classifier.fc.weight, head_b=m.Slide_classifier.fc.bias, slide_head=True,
                     shard_begin=[row_begin], group=self.group if self.exchange is None else None, exchange=self.exchange)
        return res.sub[0], res.slide, res.scores.unsqueeze(0)

    __call__ = forward
