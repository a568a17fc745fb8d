"""world_size-2 (and 3) CPU runs over gloo of the host-side sharding logic: row partition, record
all-gather, rank-0 mask draw broadcast.  The per-shard arithmetic is played by the numpy oracle here
(no GPU in this suite); the record layout is the kernels' (m[K], l[K], acc[K][L])."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gated_pool as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acmil_b200.sharding import draw_rsel, gather_records, shard_bounds
        K, Lw = 5, 16
        rng = np.random.default_rng(3)
        h = rng.standard_normal((n, Lw)).astype(np.float32)
        a = (rng.standard_normal((K, n)) * 4).astype(np.float32)
        b = shard_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
        lo, hi = b[rank], b[rank + 1]
        m, l, acc = O.pool_partials(h[lo:hi], a[:, lo:hi]) if hi > lo else (
            np.full(K, -np.inf, np.float32), np.zeros(K, np.float32), np.zeros((K, Lw), np.float32))
        rec = torch.from_numpy(np.concatenate([m, l, acc.ravel()]).astype(np.float32))
        allrec = gather_records(rec, None).numpy().reshape(world, -1)
        ms, ls, accs = [], [], []
        for r in range(world):
            if np.isinf(allrec[r, 0]):
                continue
            ms.append(allrec[r, :K]); ls.append(allrec[r, K:2 * K]); accs.append(allrec[r, 2 * K:].reshape(K, Lw))
        got, _, _ = O.merge_partials(ms, ls, accs)
        ref = O.softmax_rows(a.astype(np.float64)) @ h.astype(np.float64)
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-6)
        torch.manual_seed(100 + rank)            # ranks deliberately out of sync: the broadcast must fix it
        rsel = draw_rsel(K, 10, 6, torch.device("cpu"))
        gathered = [torch.empty_like(rsel) for _ in range(world)]
        dist.all_gather(gathered, rsel)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        assert rsel.shape == (K, 6) and int(rsel.max()) < 10
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 2)])
def test_sharded_merge_over_gloo(tmp_path, world, n):
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"ok{r}.npy") for r in range(world)]
    for o in outs[1:]:
        np.testing.assert_array_equal(outs[0], o)     # every rank finishes to the same answer
