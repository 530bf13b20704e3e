"""world_size-2 gloo tests of the ray-parallel host logic (lightplane_b200/distributed.py): shard
bounds, bucketed SUM all-reduce, and the two exchange rules -- renderer gradients add over ray
shards; splatter accumulators are reduced BEFORE normalisation -- checked with the oracle."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    from lightplane_b200.distributed import shard_bounds

    for n in (0, 1, 31, 32, 33, 1000, 2073600):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world, 32) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all((hi - lo) % 32 == 0 for lo, hi in spans if hi < n)  # only the tail is ragged
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 32 + 31


def _worker(rank, world, port, tmp):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lightplane_b200 as lp
        from lightplane_b200.distributed import (all_reduce_gradients, all_reduce_sum_, broadcast_,
                                                 shard_rays)
        from oracle import lightplane_oracle as O

        # bucketed all-reduce of several tensors
        a, b = torch.full((3, 2), float(rank + 1)), torch.arange(5.0) * (rank + 1)
        all_reduce_sum_([a, None, b])
        assert torch.allclose(a, torch.full((3, 2), 3.0)) and torch.allclose(b, torch.arange(5.0) * 3)
        w = torch.full((4,), float(rank))
        broadcast_([w], src=1)
        assert (w == 1).all()

        # renderer: gradient of the full batch == SUM over ranks of the shard gradients
        torch.manual_seed(0)
        C, H, S, n = 16, 16, 6, 96
        sizes = [[1, 1, 5, 4, C], [1, 6, 1, 4, C], [1, 6, 5, 1, C]]
        rows = sum(s[1] * s[2] * s[3] for s in sizes)
        dp = lp.init_decoder_params("cpu", 2, 2, 2, input_chn=C, hidden_chn=H, color_chn=3, opacity_init_bias=-1.0)
        o = torch.randn(n, 3) / 3
        rays = lp.Rays(directions=-o + 0.1 * torch.randn(n, 3), origins=o, grid_idx=torch.zeros(n, dtype=torch.long),
                       near=torch.full((n,), 0.1), far=torch.full((n,), 3.0), encoding=torch.randn(n, H))
        grid0 = torch.randn(rows, C)
        dims = ([C, H, H], [H, H, 1], [H, H, 16])

        def grads(r):
            grid, mlp = grid0.clone().requires_grad_(True), dp.mlp_params.clone().requires_grad_(True)
            out = O.render(r.directions, r.origins, r.grid_idx, r.near, r.far, r.encoding, grid, sizes, mlp,
                           *dims, num_samples=S, gain=1.0)
            (out[2][:, :3].sum() + out[1].sum()).backward()
            return grid, mlp

        g_full, m_full = grads(rays)
        shard = shard_rays(rays)
        assert shard.directions.shape[0] in (32, 64)
        g_loc, m_loc = grads(shard)
        all_reduce_gradients([g_loc, m_loc])
        assert torch.allclose(g_loc.grad, g_full.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(m_loc.grad, m_full.grad, rtol=1e-4, atol=1e-5)

        # splatter: reduce un-normalised feature and weight accumulators, THEN normalise
        feat = torch.rand(n, C)
        out_sizes = [[1, 6, 5, 4, C]]

        def accumulators(r, f):
            d, _ = O.ray_depths(r.near, r.far, S, 0, 1e-5)
            pts = r.origins[:, None] + d[..., None] * r.directions[:, None]
            acc = O.splat_grid_list(torch.zeros(120, C), out_sizes, r.grid_idx, pts, f[:, None].expand(-1, S, -1), False)
            wgt = O.splat_grid_list(torch.zeros(120, 1), [[1, 6, 5, 4, 1]], r.grid_idx, pts, torch.ones(len(f), S, 1), False)
            return acc, wgt

        full = O.splat(rays.directions, rays.origins, rays.grid_idx, rays.near, rays.far, feat, out_sizes, num_samples=S)
        lo = 0 if rank == 0 else 64
        acc, wgt = accumulators(shard, feat[lo: lo + shard.directions.shape[0]])
        all_reduce_sum_([acc, wgt])
        assert torch.allclose(acc / wgt.clamp(min=1e-5), full, rtol=1e-4, atol=1e-5)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
