#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, NCCL): rays sharded over ranks + gradient /
accumulator all-reduce must equal the single-GPU result on the full batch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import lightplane_b200 as lp  # noqa: E402
from lightplane_b200.distributed import all_reduce_gradients, broadcast_, shard_rays  # noqa: E402


def rel(a, b):
    return float((a - b).abs().mean() / b.abs().mean())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)  # identical data on every rank
    C, H, S, n = 16, 32, 64, 8192
    shapes = [[1, 1, 32, 32, C], [1, 32, 1, 32, C], [1, 32, 32, 1, C]]
    dp = lp.init_decoder_params(dev, 2, 2, 2, input_chn=C, hidden_chn=H, color_chn=3, opacity_init_bias=-1.0)
    grids = [torch.randn(s, device=dev) for s in shapes]
    broadcast_([dp.mlp_params] + grids, src=0)
    o = torch.randn(n, 3, device=dev) / 3
    rays = lp.Rays(directions=-o + 0.1 * torch.randn(n, 3, device=dev), origins=o,
                   grid_idx=torch.zeros(n, dtype=torch.int64, device=dev), near=torch.full((n,), 0.1, device=dev),
                   far=torch.full((n,), 3.0, device=dev), encoding=torch.randn(n, H, device=dev))

    def grads(r):
        g = [x.clone().requires_grad_(True) for x in grids]
        m = dp.mlp_params.clone().requires_grad_(True)
        out = lp.lightplane_renderer(r, g, lp.DecoderParams(m, dp.n_hidden_trunk, dp.n_hidden_opacity, dp.n_hidden_color, 3),
                                     num_samples=S, gain=1.0)
        (out[2].sum() + out[1].sum()).backward()
        return g + [m]

    full = grads(rays)
    local = grads(shard_rays(rays))
    all_reduce_gradients(local)
    errs = [rel(a.grad, b.grad) for a, b in zip(local, full)]
    feat = torch.rand(n, C, device=dev)
    srays = lp.Rays(directions=rays.directions, origins=rays.origins, grid_idx=rays.grid_idx, near=rays.near,
                    far=rays.far, encoding=feat)
    sizes = [(1, 24, 24, 24, C)]
    full_s = lp.lightplane_splatter(srays, sizes, num_samples=S, return_list=False)
    shard_s = lp.lightplane_splatter(shard_rays(srays), sizes, num_samples=S, return_list=False,
                                     process_group=dist.group.WORLD)
    errs.append(rel(shard_s, full_s))
    ok = all(e < 2e-3 for e in errs)
    print(f"rank {rank}/{world}: sharded-vs-full rel errors {['%.1e' % e for e in errs]} -> {'OK' if ok else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
