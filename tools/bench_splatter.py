#!/usr/bin/env python
"""Splatter benchmark (BASELINE.json configs[3]): 100 views x 256x256 rays, 256 samples into a
128^3 x 32-channel voxel grid, forward + backward through the public `lightplane_splatter` op.
Prints one JSON line: rays/s, per-kernel times, scattered-bytes/s against the measured HBM peak."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import lightplane_b200 as lp  # noqa: E402
from lightplane_b200 import _cabi  # noqa: E402
from bench import camera_rays, load_peaks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=100)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--chn", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda")
    parts = [camera_rays(a.res, a.res, 100 + v, "cpu") for v in range(a.views)]
    d, o, gi, nr, fr = (torch.cat([p[i] for p in parts]).to(dev) for i in range(5))
    n = d.shape[0]
    feat = torch.rand(n, a.chn, device=dev, requires_grad=True)
    rays = lp.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=feat)
    sizes = [(1, a.grid, a.grid, a.grid, a.chn)]
    cot = torch.randn(a.grid ** 3, a.chn, device=dev)

    def step():
        feat.grad = None
        out = lp.lightplane_splatter(rays, sizes, num_samples=a.samples, return_list=False)
        (out * cot).sum().backward()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    _cabi.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    launches = _cabi.profile_end()
    ms = e0.elapsed_time(e1) / a.steps
    per = {}
    for k, t in launches:
        per.setdefault(k, []).append(t)
    avg = {k: sum(v) / len(v) for k, v in per.items()}
    samples = n * a.samples
    peaks = load_peaks()
    fwd_bytes = samples * 8 * (a.chn * 4 + 4)   # 8 taps x (C floats + 1 weight) of atomic traffic
    bwd_bytes = samples * 8 * a.chn * 4
    print(json.dumps({
        "metric": "splatter_fwd_bwd_rays_per_s", "value": n / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
        "config": {"rays": n, "samples": a.samples, "grid": f"{a.grid}^3x{a.chn}"},
        "launch_ms": avg,
        "fwd_scatter_gbs": fwd_bytes / (avg["lp_splat_forward"] * 1e-3) / 1e9,
        "bwd_gather_gbs": bwd_bytes / (avg["lp_splat_backward"] * 1e-3) / 1e9,
        "hbm_peak_gbs": peaks["hbm_gbs"],
    }))


if __name__ == "__main__":
    main()
