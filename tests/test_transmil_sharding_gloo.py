"""Host logic of the sequence-parallel TransMIL (acmil_b200/transmil_sharded.py) on CPU: the token partition, and a
world_size-2 / 4 run over gloo in which every rank plays its share of one Nystrom layer in numpy -- same partition, same
exchanges through DistComm (landmark all-gather, pseudo-inverse heads, log-sum-exp merge of the attn3 v partial sums, conv
halo) -- against the whole-sequence oracle (oracle/transmil.py, pinned to reference fixtures).  The CUDA kernels replace
the numpy arithmetic on the GPU box (tests/test_transmil_gpu.py runs the same per-rank code with ThreadComm)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import transmil as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n,world,m", [(50000, 8, 256), (50000, 2, 256), (1000, 4, 64), (300, 2, 32), (97, 4, 16)])
def test_shard_plan_partitions_the_token_sequence(n, world, m):
    from acmil_b200.transmil_sharded import ShardPlan
    p = ShardPlan(n, world, m)
    g = int(np.ceil(np.sqrt(n)))
    assert p.T == 1 + g * g and p.n_pad % m == 0 and p.n_pad - p.T == p.pad and 0 <= p.pad < m
    assert p.U * world == p.n_pad and p.U == p.m_loc * p.l                      # cuts fall on landmark-group boundaries
    toks = [p.tokens(r) for r in range(world)]
    assert toks[0][0] == 0 and toks[-1][1] == p.T and all(toks[r][1] == toks[r + 1][0] for r in range(world - 1))
    assert all(t1 - t0 == p.U - p.lead_zero(r) for r, (t0, t1) in enumerate(toks))
    # the patch rows behind the tokens: [cls, patches, the first `add` patches again]  (transMIL.py:63-72)
    seq = np.concatenate([np.arange(n), np.arange(g * g - n)])
    got = np.concatenate([p.patch_rows(r) for r in range(world)])
    np.testing.assert_array_equal(got, seq)


def test_shard_plan_rejects_what_cannot_be_cut():
    from acmil_b200.transmil_sharded import ShardPlan
    with pytest.raises(ValueError):
        ShardPlan(1000, 3, 64)          # 64 landmarks over 3 ranks
    with pytest.raises(ValueError):
        ShardPlan(20, 8, 16)            # front padding larger than a rank's share


def _layer_shard_numpy(p, x_loc, plan, rank, comm, heads, d):
    """one rank's share of NystromAttention (no pre-norm, no outer residual) in numpy, exchanges through `comm`"""
    f = np.float64
    m, l, lead = plan.m, plan.l, plan.lead_zero(rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))      # noqa: E731
    wqkv = p["to_qkv.weight"].astype(f)
    xp = np.concatenate([np.zeros((lead, x_loc.shape[1]), f), x_loc.astype(f)])
    q, k, v = np.split(xp @ wqkv.T, 3, axis=-1)
    hd = lambda a: a.reshape(-1, heads, d).transpose(1, 0, 2)    # noqa: E731
    q, k, v = hd(q) * d ** -0.5, hd(k), hd(v)
    ql_loc = q.reshape(heads, plan.m_loc, l, d).mean(2)
    kl_loc = k.reshape(heads, plan.m_loc, l, d).mean(2)
    ql = torch.cat(comm.all_gather(rank, t(ql_loc)), 1).numpy()
    kl = torch.cat(comm.all_gather(rank, t(kl_loc)), 1).numpy()
    a2 = O._softmax(ql @ kl.transpose(0, 2, 1))
    hc = heads // comm.world
    hf = rank * hc
    # pseudo-inverse of the own heads; the start value's scale is a maximum over ALL heads
    ax = np.abs(a2)
    scale = ax.sum(-1).max() * ax.sum(-2).max()
    z = a2[hf:hf + hc].transpose(0, 2, 1) / scale
    eye = np.eye(m)
    for _ in range(6):
        xz = a2[hf:hf + hc] @ z
        z = 0.25 * z @ (13 * eye - (xz @ (15 * eye - (xz @ (7 * eye - xz)))))
    z = torch.cat(comm.all_gather(rank, t(z)), 0).numpy()
    s3 = ql @ k.transpose(0, 2, 1)                                # [heads, m, n_loc]
    mx = s3.max(-1)
    e = np.exp(s3 - mx[..., None])
    parts = torch.stack(comm.all_gather(rank, t(e @ v))).numpy()  # [P, heads, m, d]
    ms = torch.stack(comm.all_gather(rank, t(mx))).numpy()
    ls = torch.stack(comm.all_gather(rank, t(e.sum(-1)))).numpy()
    M = ms.max(0)
    wgt = np.exp(ms - M)
    kv = (parts * wgt[..., None]).sum(0) / (wgt * ls).sum(0)[..., None]
    out = (O._softmax(q @ kl.transpose(0, 2, 1)) @ z) @ kv       # [heads, n_loc, d]
    wc = p["res_conv.weight"].astype(f)
    half = wc.shape[2] // 2
    fl, fr = comm.shift(rank, t(v[:, :half]), t(v[:, -half:]))
    zero = np.zeros((heads, half, d))
    vp = np.concatenate([zero if fl is None else fl.numpy(), v, zero if fr is None else fr.numpy()], 1)
    for tt in range(wc.shape[2]):
        out = out + wc[:, 0, tt, 0, None, None] * vp[:, tt:tt + v.shape[1]]
    out = out.transpose(1, 0, 2).reshape(-1, heads * d) @ p["to_out.0.weight"].astype(f).T + p["to_out.0.bias"].astype(f)
    return out[lead:]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acmil_b200.transmil_sharded import DistComm, ShardPlan
        heads, d, m, dim, n = 8, 4, 16, 32, 200
        rng = np.random.default_rng(5)
        p = {"to_qkv.weight": rng.standard_normal((3 * heads * d, dim)) * 0.3, "to_out.0.weight": rng.standard_normal((dim, heads * d)) * 0.2,
             "to_out.0.bias": rng.standard_normal(dim) * 0.1, "res_conv.weight": rng.standard_normal((heads, 1, 9, 1)) * 0.2}
        plan = ShardPlan(n, world, m)
        x = rng.standard_normal((plan.T, dim))                    # the token sequence of one layer
        t0, t1 = plan.tokens(rank)
        got = _layer_shard_numpy(p, x[t0:t1], plan, rank, DistComm(), heads, d)
        ref = O.nystrom_attention(p, x[None], heads=heads, dim_head=d, num_landmarks=m, dtype=np.float64)[0]
        np.testing.assert_allclose(got, ref[t0:t1], rtol=1e-8, atol=1e-9)
        # broadcast0: every rank ends with rank 0's tensor
        v = DistComm().broadcast0(rank, torch.full((3,), float(rank)))
        assert float(v.sum()) == 0.0
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nystrom_layer_over_gloo(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(world))
