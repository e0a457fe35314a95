"""The N > 1 host path on CPU: chain sharding and the end-of-run all-gather of draws over gloo (world_size 2),
plus the R-hat / ESS diagnostics that consume the gathered draws."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, tmp):
    import torch
    import torch.distributed as dist
    from manifold_constrained_gaussian_process_inference_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n_local = D.shard_chains(n_total, rank, world)
    # draws[it, c, col] = 1000*it + global chain id + 0.1*col : lets the test verify global ordering after the gather
    it = torch.arange(5, dtype=torch.float64)[:, None, None]
    ch = (first + torch.arange(n_local, dtype=torch.float64))[None, :, None]
    col = torch.arange(3, dtype=torch.float64)[None, None, :]
    local = 1000 * it + ch + 0.1 * col
    full = D.allgather_draws(local)
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), full.numpy())
    dist.destroy_process_group()


def test_shard_chains_partition():
    from manifold_constrained_gaussian_process_inference_b200 import distributed as D
    for total in (1, 7, 8, 65536, 4097):
        for world in (1, 2, 3, 8):
            parts = [D.shard_chains(total, r, world) for r in range(world)]
            assert sum(n for _, n in parts) == total
            assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1


def test_allgather_draws_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 7, str(tmp_path)), nprocs=2, join=True)       # 7 chains: ragged shards (4 + 3)
    g = np.load(tmp_path / "gathered.npy")
    assert g.shape == (5, 7, 3)
    it, ch, col = np.meshgrid(np.arange(5), np.arange(7), np.arange(3), indexing="ij")
    assert np.allclose(g, 1000 * it + ch + 0.1 * col)


def test_rhat_and_ess_on_known_processes():
    from manifold_constrained_gaussian_process_inference_b200 import diagnostics as dg
    rng = np.random.default_rng(0)
    iid = rng.normal(size=(1000, 8))
    assert abs(dg.split_rhat(iid) - 1.0) < 0.02
    ess = dg.ess_bulk(iid)
    assert 5000 < ess < 11000                                   # ~ n * m for independent draws
    shifted = iid + np.arange(8)[None, :]                       # chains with different means must be flagged
    assert dg.split_rhat(shifted) > 1.5
    phi = 0.9                                                   # AR(1): ESS ~ N (1 - phi) / (1 + phi)
    ar = np.zeros((4000, 4)); e = rng.normal(size=ar.shape)
    for i in range(1, 4000):
        ar[i] = phi * ar[i - 1] + e[i]
    ess_ar = dg.ess_bulk(ar)
    assert 0.5 * 16000 * (1 - phi) / (1 + phi) < ess_ar < 2.0 * 16000 * (1 - phi) / (1 + phi)
    s = dg.summarize(np.stack([iid, shifted], axis=2), names=["ok", "bad"])
    assert s[0]["rhat"] < 1.05 < s[1]["rhat"]
