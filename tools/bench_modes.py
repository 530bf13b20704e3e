#!/usr/bin/env python
"""Timing of the renderer configurations that take the generic kernels and of the MLP
splatter (8f rank 1), forward + backward through the public ops; one JSON line per mode."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import lightplane_b200 as lp  # noqa: E402
from bench import camera_rays  # noqa: E402


def timed(fn, steps=3, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--samples", type=int, default=128)
    a = ap.parse_args()
    dev = torch.device("cuda")
    d, o, gi, nr, fr = (t.to(dev) for t in camera_rays(a.res, a.res, 0, "cpu"))
    n = d.shape[0]
    C, P = 16, 64
    shapes = [[1, 1, P, P, C], [1, P, 1, P, C], [1, P, P, 1, C]]

    def renderer(tag, layers, color_grid=False, scaffold=False, hidden=32, C=C, P=P):
        nt, no, nc = layers
        shapes = [[1, 1, P, P, C], [1, P, 1, P, C], [1, P, P, 1, C]]
        dp = lp.init_decoder_params(dev, no, nt, nc, input_chn=C, hidden_chn=hidden, color_chn=3, opacity_init_bias=-1.0,
                                    use_separate_color_grid=color_grid)
        dp.mlp_params.requires_grad_(True)
        grids = [(0.5 * torch.randn(s, device=dev)).requires_grad_(True) for s in shapes]
        cgrids = [(0.5 * torch.randn(s, device=dev)).requires_grad_(True) for s in shapes] if color_grid else None
        enc = torch.randn(n, C if color_grid else hidden, device=dev, requires_grad=True)
        scaf = (torch.rand(1, 32, 32, 32, device=dev) > 0.5).float() if scaffold else None
        rays = lp.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=enc)
        tgt = torch.rand(n, 3, device=dev)

        def step():
            for t in grids + [dp.mlp_params, enc] + (cgrids or []):
                t.grad = None
            _, _, f = lp.lightplane_renderer(rays, grids, dp, num_samples=a.samples, gain=1.0, color_grid=cgrids, scaffold=scaf)
            ((f - tgt) ** 2).sum().backward()

        def fwd_only():
            with torch.no_grad():
                lp.lightplane_renderer(rays, grids, dp, num_samples=a.samples, gain=1.0, color_grid=cgrids, scaffold=scaf)

        ms, ms_f = timed(step), timed(fwd_only)
        print(json.dumps({"mode": tag, "rays": n, "samples": a.samples, "ms_fwd_bwd": ms, "rays_per_s": n / ms * 1e3,
                          "ms_fwd_only": ms_f, "fwd_rays_per_s": n / ms_f * 1e3}))

    renderer("renderer 2/2/2 (tensor-core path)", (2, 2, 2))
    renderer("renderer 2/2/2 + scaffold (tensor-core path)", (2, 2, 2), scaffold=True)
    renderer("renderer 0/2/2 colour grid (tensor-core path)", (0, 2, 2), color_grid=True)
    renderer("renderer 4/2/4 (tensor-core path, table driven)", (4, 2, 4))
    renderer("renderer 4/4/4 (generic)", (4, 4, 4))
    renderer("renderer 2/2/2 hidden 64, 128^2x32 planes, scaffold = the reference's example config (tensor-core path)", (2, 2, 2),
             scaffold=True, hidden=64, C=32, P=128)

    # MLP splatter: 32^3 x 16 input grid -> 64^3 x 16 output grid
    in_sizes, out_sizes = [(1, 32, 32, 32, 16)], [(1, 64, 64, 64, 16)]
    ing = [(0.5 * torch.randn(s, device=dev)).requires_grad_(True) for s in in_sizes]
    mlp = lp.init_splatter_params(dev, 2, input_chn=16, hidden_chn=32, out_chn=16) if hasattr(lp, "init_splatter_params") else None
    if mlp is not None:
        mlp.mlp_params.requires_grad_(True)
        feat = torch.rand(n, 16, device=dev, requires_grad=True)
        rays = lp.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=feat)
        cot = torch.randn(64 ** 3, 16, device=dev)

        def step2():
            out = lp.lightplane_mlp_splatter(rays, out_sizes, mlp, ing, num_samples=a.samples, return_list=False)
            (out * cot).sum().backward()

        ms = timed(step2)
        print(json.dumps({"mode": "mlp splatter 2 layers (tensor-core path)", "rays": n, "samples": a.samples, "ms_fwd_bwd": ms,
                          "rays_per_s": n / ms * 1e3}))


if __name__ == "__main__":
    main()
