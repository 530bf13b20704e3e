"""Oracle comparisons AT the BASELINE.json configurations (VERDICT r1 "weak" #1): the bench camera,
128 samples, the 64^2 x 16 triplane with Xavier-scale decoder weights -- every output and every gradient
against the fp64 oracle, for both upstream losses of SURVEY.md 8d; and the splatter's configs[3] shape
(256 samples into a 128^3 x 32 voxel grid).  Ray counts are what the CPU oracle finishes in seconds;
the full-size runs are covered by the size-independent properties of test_gpu_parity.py and by
`bench.py`'s `parity_err` (GPU vs oracle on the bench workload's own rays)."""
import os
import sys

import pytest
import torch

from _golden import rel_err

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu
TOL_OUT, TOL_GRAD = 2e-4, 1e-3   # north_star: 1e-3 on features / alpha / grid gradients
TOL_GMLP = 3e-3                  # parameter gradients: reduced from bf16 operand tiles; measured 1.2-1.5e-3 here with
                                 # the random-sign loss, 2e-4 with the image loss (the reference's own criterion is 7e-4 mean-rel)


def _bench_problem(side, seed, dev):
    import lightplane_b200 as lp
    from bench import camera_rays

    torch.manual_seed(seed)
    C, H = 16, 32
    dp = lp.init_decoder_params(dev, 2, 2, 2, input_chn=C, hidden_chn=H, color_chn=3, opacity_init_bias=-1.0)
    shapes = [[1, 1, 64, 64, C], [1, 64, 1, 64, C], [1, 64, 64, 1, C]]
    grids = [0.5 * torch.randn(s, device=dev) for s in shapes]
    d, o, gi, nr, fr = [t.to(dev) for t in camera_rays(side, side, 1000 + seed, "cpu")]
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(side * side, H, generator=g).to(dev)
    return dp, shapes, grids, (d, o, gi, nr, fr), enc


@pytest.mark.parametrize("loss_kind,tile_walk", [("randsign", False), ("mse", False), ("randsign", True)])
def test_renderer_bench_camera_4096x128_vs_oracle(loss_kind, tile_walk):
    """configs[1]/[2] workload shape (bench camera, 128 samples, 64^2x16 triplane, 2/2/2 h32) on 64x64 rays."""
    errs = bench_camera_errors(loss_kind, tile_walk)
    print(f"bench-camera 4096x128 [{loss_kind}, tile_walk={tile_walk}] errors vs fp64 oracle:", {k: f"{v:.1e}" for k, v in errs.items()})
    for k in ("ray_length", "nlt", "features"):
        assert errs[k] < TOL_OUT, errs
    assert errs["g_grid"] < TOL_GRAD and errs["g_enc"] < TOL_GRAD, errs
    assert errs["g_mlp"] < TOL_GMLP, errs


def bench_camera_errors(loss_kind, tile_walk):
    """Relative errors of every output and gradient against the fp64 oracle (also used by tools/fp32_floor.py)."""
    import lightplane_b200 as lp
    from oracle import lightplane_oracle as O

    dev, side, S, C, H = "cuda", 64, 128, 16, 32
    dp, shapes, grids, (d, o, gi, nr, fr), enc = _bench_problem(side, 0, dev)
    n = side * side
    g = torch.Generator().manual_seed(11)
    cot = [torch.randn(n, generator=g).to(dev), torch.randn(n, generator=g).to(dev), torch.randn(n, 3, generator=g).to(dev)]
    target = torch.rand(n, 3, generator=g).to(dev)

    def loss_fn(outs, f=lambda t: t):
        if loss_kind == "randsign":   # the reference tests' loss (tests/test_renderer_with_autograd.py:211-213)
            return sum((f(c) * v).sum() for c, v in zip(cot, outs))
        return ((outs[2] - f(target)) ** 2).mean()   # image MSE (the bench's loss)

    gl = [x.clone().requires_grad_(True) for x in grids]
    mp = dp.mlp_params.detach().clone().requires_grad_(True)
    e = enc.clone().requires_grad_(True)
    dpp = lp.DecoderParams(mp, dp.n_hidden_trunk, dp.n_hidden_opacity, dp.n_hidden_color, dp.color_chn)
    rays = lp.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=e)
    outs = lp.lightplane_renderer(rays, gl, dpp, num_samples=S, gain=1.0, ray_image_width=side if tile_walk else None)
    grads = torch.autograd.grad(loss_fn(outs), gl + [mp, e])

    f = lambda t: t.detach().double().cpu()
    og = f(torch.cat([x.reshape(-1, C) for x in grids], 0)).requires_grad_(True)
    om, oe = f(dp.mlp_params).requires_grad_(True), f(enc).requires_grad_(True)
    oo = O.render(f(d), f(o), gi.cpu().long(), f(nr), f(fr), oe, og, shapes, om, [C, H, H], [H, H, 1], [H, H, 16],
                  num_samples=S, gain=1.0)
    oo = (oo[0], oo[1], oo[2][:, :3])
    ograds = torch.autograd.grad(loss_fn(oo, f), [og, om, oe])
    errs = {nm: rel_err(a, b) for a, b, nm in zip(outs, oo, ("ray_length", "nlt", "features"))}
    errs["g_grid"] = rel_err(torch.cat([x.reshape(-1, C) for x in grads[:3]], 0), ograds[0])
    errs["g_mlp"] = rel_err(grads[3], ograds[1])
    errs["g_enc"] = rel_err(grads[4], ograds[2])
    return errs


def test_splatter_cfg4_shape_8192x256_vs_oracle():
    """configs[3] shape: 256 samples into a [1,128,128,128,32] voxel grid, 8192 camera rays, forward + backward."""
    import lightplane_b200 as lp
    from bench import camera_rays
    from oracle import lightplane_oracle as O

    dev, C, S = "cuda", 32, 256
    sizes = [[1, 128, 128, 128, C]]
    d, o, gi, nr, fr = [t.to(dev) for t in camera_rays(128, 64, 1003, "cpu")]
    n = d.shape[0]
    g = torch.Generator().manual_seed(5)
    feat = torch.rand(n, C, generator=g).to(dev).requires_grad_(True)
    rays = lp.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=feat)
    out = lp.lightplane_splatter(rays, [tuple(s) for s in sizes], num_samples=S, return_list=False)
    cot = torch.randn(out.shape, generator=g).to(dev)
    (g_feat,) = torch.autograd.grad((out * cot).sum(), [feat])
    f = lambda t: t.detach().double().cpu()
    ofeat = f(feat).requires_grad_(True)
    oout = O.splat(f(d), f(o), gi.cpu().long(), f(nr), f(fr), ofeat, sizes, num_samples=S)
    (og,) = torch.autograd.grad((oout * f(cot)).sum(), [ofeat])
    errs = dict(out=rel_err(out, oout), g_feat=rel_err(g_feat, og))
    print("splatter 8192x256 -> 128^3x32 errors vs fp64 oracle:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert errs["out"] < TOL_OUT and errs["g_feat"] < TOL_OUT, errs
