#!/usr/bin/env python
"""Splatter benchmark (BASELINE.json configs[3]): 100 views x 256x256 rays, 256 samples into a
128^3 x 32-channel voxel grid, forward + backward through the public `lightplane_splatter` op.
Prints one JSON line: rays/s, per-kernel times, and the reduction / gather bytes the kernels ACTUALLY issue (taps with
a non-zero weight: `lp_splat_fwd_kernel` skips the others, and most samples of a scene view lie outside the grid)
per second against the measured HBM peak.  DRAM / L2-reduction counters of the same kernels are taken with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum \
      -k regex:lp_splat --csv --log-file gpurun_out/splat_ncu.csv python tools/bench_splatter.py --steps 1 --warmup 1
(profiles/splatter_r2.md)."""
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


def count_taps(d, o, near, far, samples, grid):
    """Number of (sample, corner) pairs with a non-zero trilinear weight inside a grid^3 voxel grid (what the forward
    kernel reduces and the backward kernel gathers), and the share of samples that touch the grid at all."""
    frac = torch.linspace(0, 1, samples, device=d.device)
    taps = touched = 0
    for lo in range(0, d.shape[0], 1 << 16):
        sl = slice(lo, lo + (1 << 16))
        t = near[sl, None] + (far[sl] - near[sl])[:, None] * frac[None]
        p = o[sl, None, :] + t[..., None] * d[sl, None, :]
        i = (p + 1) * 0.5 * grid - 0.5
        i0 = torch.floor(i)
        f = i - i0
        n0 = ((i0 >= 0) & (i0 < grid) & (f < 1)).long()          # lower corner inside with weight 1-f > 0
        n1 = ((i0 + 1 >= 0) & (i0 + 1 < grid) & (f > 0)).long()  # upper corner inside with weight f > 0
        per_axis = n0 + n1
        cnt = per_axis[..., 0] * per_axis[..., 1] * per_axis[..., 2]
        taps += int(cnt.sum())
        touched += int((cnt > 0).sum())
    return taps, touched


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
    taps, touched = count_taps(d, o, nr, fr, a.samples, a.grid)
    fwd_bytes = taps * (a.chn * 4 + 4)   # per non-zero tap: C floats + 1 weight of fp32 reductions
    bwd_bytes = taps * a.chn * 4         # per non-zero tap: a C-float row gathered
    grid_bytes = a.grid ** 3 * (a.chn + 1) * 4
    print(json.dumps({
        "metric": "splatter_fwd_bwd_rays_per_s", "value": n / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
        "config": {"rays": n, "samples": a.samples, "grid": f"{a.grid}^3x{a.chn}"},
        "launch_ms": avg,
        "samples_touching_grid_frac": touched / samples, "nonzero_taps_per_sample": taps / samples,
        "fwd_reduced_bytes": fwd_bytes, "bwd_gathered_bytes": bwd_bytes, "grid_bytes": grid_bytes,
        "fwd_scatter_gbs": fwd_bytes / (avg["lp_splat_forward"] * 1e-3) / 1e9,
        "bwd_gather_gbs": bwd_bytes / (avg["lp_splat_backward"] * 1e-3) / 1e9,
        "hbm_peak_gbs": peaks["hbm_gbs"],
        "fwd_frac_of_hbm_peak": fwd_bytes / (avg["lp_splat_forward"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
        "bwd_frac_of_hbm_peak": bwd_bytes / (avg["lp_splat_backward"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
        "note": "the 256 MiB feature grid + weight grid exceed the 126 MB L2: reductions are read-modify-writes at DRAM "
                "(ncu dram bytes in profiles/splatter_r2.md); bytes counted are those of taps with non-zero weight only",
    }))


if __name__ == "__main__":
    main()
