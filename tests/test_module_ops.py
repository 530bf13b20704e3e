"""The renderer module's fused glue (csrc/lp_ray_embed.cuh; reference renderer_module.py:552-601,
ray_utils.py:181-212) against the PyTorch composition of the same formulas: through the C-ABI of the
host-emulated build here, and through the public module on the GPU."""
import os
import subprocess

import pytest
import torch

from _golden import rel_err
from lightplane_b200 import _cabi
from lightplane_b200.ray_utils import calc_harmonic_embedding

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "lightplane_b200", "csrc")


def _problem(n, harm, out, seed=0, dev="cpu"):
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g) * torch.rand(n, 1, generator=g) * 3
    d[0] = 0  # a degenerate direction: F.normalize clamps the norm at 1e-12
    w = torch.randn(out, 3 + 6 * harm, generator=g) * 0.3
    b = torch.randn(out, generator=g) * 0.1
    go = torch.randn(n, out, generator=g)
    return [t.to(dev) for t in (d, w, b, go)]


def _torch_embed(d, w, b, harm):
    e = calc_harmonic_embedding(torch.nn.functional.normalize(d, dim=-1), harm)
    return torch.nn.functional.linear(e, w, b)


def _cabi_embed(lib, d, w, b, go, harm, dev):
    n, out = d.shape[0], w.shape[0]
    st = _cabi.stream_ptr(torch.device(dev))
    enc = torch.empty(n, out, device=dev)
    _cabi.check(lib, lib.lp_ray_embed_forward(st, n, d.data_ptr(), harm, w.data_ptr(), b.data_ptr(), out, enc.data_ptr()), "fwd")
    gw, gb = torch.zeros_like(w), torch.zeros_like(b)
    _cabi.check(lib, lib.lp_ray_embed_backward(st, n, d.data_ptr(), harm, go.data_ptr(), out, gw.data_ptr(), gb.data_ptr()), "bwd")
    if dev != "cpu":
        torch.cuda.synchronize()
    return enc, gw, gb


def _check_embed(lib, n, harm, out, dev):
    d, w, b, go = _problem(n, harm, out, dev=dev)
    enc, gw, gb = _cabi_embed(lib, d, w, b, go, harm, dev)
    dd, wd, bd, gd = (t.double().cpu() for t in (d, w, b, go))
    wd.requires_grad_(True)
    bd.requires_grad_(True)
    want = _torch_embed(dd, wd, bd, harm)
    gww, gbw = torch.autograd.grad((want * gd).sum(), [wd, bd])
    # fp32 rounding of the unit direction is amplified by the top frequency 2^(harm-1) inside sin(): the PyTorch fp32
    # composition differs from fp64 by the same amount
    tol = 1e-6 * (2 + 2 ** harm)
    assert rel_err(enc, want.detach()) < tol, rel_err(enc, want.detach())
    assert rel_err(gw, gww) < 10 * tol and rel_err(gb, gbw) < 2e-5, (rel_err(gw, gww), rel_err(gb, gbw))


def _check_bg(lib, n, c, log_t, dev):
    g = torch.Generator().manual_seed(3)
    nlt, feat = (torch.rand(n, generator=g) * 4).to(dev), torch.randn(n, c, generator=g).to(dev)
    bg, ga, go = torch.rand(c, generator=g).to(dev), torch.randn(n, generator=g).to(dev), torch.randn(n, c, generator=g).to(dev)
    st = _cabi.stream_ptr(torch.device(dev))
    alpha, out, gn = torch.empty_like(nlt), torch.empty_like(feat), torch.empty_like(nlt)
    _cabi.check(lib, lib.lp_bg_composite_forward(st, n, c, nlt.data_ptr(), feat.data_ptr(), bg.data_ptr(), int(log_t),
                                                 alpha.data_ptr(), out.data_ptr()), "bg fwd")
    _cabi.check(lib, lib.lp_bg_composite_backward(st, n, c, nlt.data_ptr(), bg.data_ptr(), int(log_t), ga.data_ptr(),
                                                  go.data_ptr(), gn.data_ptr()), "bg bwd")
    if dev != "cpu":
        torch.cuda.synchronize()
    l = nlt.double().cpu().requires_grad_(True)
    T = torch.exp(-l)
    wo = feat.double().cpu() + T[:, None] * bg.double().cpu()
    wa = -l if log_t else 1 - T
    (gl,) = torch.autograd.grad((wa * ga.double().cpu()).sum() + (wo * go.double().cpu()).sum(), [l])
    assert rel_err(out, wo.detach()) < 1e-6 and rel_err(alpha, wa.detach()) < 1e-6 and rel_err(gn, gl) < 2e-6


@pytest.fixture(scope="module")
def hostlib():
    subprocess.run(["make", "-s", "-C", CSRC, "hostsim"], check=True)
    return _cabi.load_library(os.path.join(HERE, "hostsim", "liblp_hostsim.so"))


@pytest.mark.parametrize("n,harm,out", [(1, 3, 32), (257, 3, 32), (700, 0, 4), (300, 10, 16), (513, 2, 64)])
def test_hostsim_ray_embed(hostlib, n, harm, out):
    _check_embed(hostlib, n, harm, out, "cpu")


@pytest.mark.parametrize("log_t", [False, True])
def test_hostsim_bg_composite(hostlib, log_t):
    _check_bg(hostlib, 777, 3, log_t, "cpu")


def test_ray_embed_rejects_unsupported_shapes(hostlib):
    d, w, b, go = _problem(8, 3, 32)
    enc = torch.empty(8, 32)
    assert hostlib.lp_ray_embed_forward(None, 8, d.data_ptr(), 11, w.data_ptr(), b.data_ptr(), 32, enc.data_ptr()) != 0
    assert hostlib.lp_ray_embed_forward(None, 8, d.data_ptr(), 3, w.data_ptr(), b.data_ptr(), 30, enc.data_ptr()) != 0
    assert hostlib.lp_ray_embed_forward(None, 8, None, 3, w.data_ptr(), b.data_ptr(), 32, enc.data_ptr()) != 0
    assert hostlib.lp_ray_embed_forward(None, 0, None, 3, w.data_ptr(), b.data_ptr(), 32, enc.data_ptr()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,harm,out", [(1, 3, 32), (100003, 3, 32), (4097, 0, 4), (5000, 10, 16), (70001, 2, 64)])
def test_gpu_ray_embed(n, harm, out):
    _check_embed(_cabi.get_lib(), n, harm, out, "cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("log_t", [False, True])
def test_gpu_bg_composite(log_t):
    _check_bg(_cabi.get_lib(), 100001, 3, log_t, "cuda")


@pytest.mark.gpu
def test_gpu_module_uses_fused_glue_and_matches_composition():
    """`LightplaneRenderer.forward`: same outputs and parameter gradients with the fused glue as with the PyTorch
    composition (forced by a direction tensor that requires a gradient and a bg colour that does)."""
    import lightplane_b200 as lp

    torch.manual_seed(0)
    dev, n = "cuda", 2048
    m = lp.LightplaneRenderer(num_samples=24, color_chn=3, grid_chn=16, mlp_hidden_chn=32, bg_color=(0.2, 0.5, 0.9),
                              opacity_init_bias=-1.0).to(dev)
    grids = [(0.5 * torch.randn(1, *s, 16, device=dev)).requires_grad_(True) for s in ((1, 16, 16), (16, 1, 16), (16, 16, 1))]
    d = torch.nn.functional.normalize(torch.randn(n, 3, device=dev) * 0.2 + torch.tensor([0.0, 0.0, 1.0], device=dev), dim=-1)
    o = torch.tensor([0.0, 0.0, -2.0], device=dev).expand(n, 3).contiguous()
    mk = lambda dd: lp.Rays(directions=dd, origins=o, grid_idx=torch.zeros(n, dtype=torch.int32, device=dev),
                            near=torch.full((n,), 1.0, device=dev), far=torch.full((n,), 3.0, device=dev))
    params = list(m.parameters()) + grids
    cot = torch.randn(n, 3, device=dev)

    def run(dirs, bg):
        _cabi.profile_begin()
        rl, alpha, feat = m(mk(dirs), grids, bg_color=bg)
        gr = torch.autograd.grad((feat * cot).sum() + alpha.sum() + rl.sum(), params)
        names = [k for k, _ in _cabi.profile_end()]
        return (rl, alpha, feat), gr, names

    outs_f, grads_f, names_f = run(d, None)
    outs_t, grads_t, names_t = run(d.clone().requires_grad_(True), torch.tensor([0.2, 0.5, 0.9], device=dev, requires_grad=True))
    assert "lp_ray_embed_forward" in names_f and "lp_ray_embed_backward" in names_f and "lp_bg_composite_forward" in names_f
    assert "lp_ray_embed_forward" not in names_t and "lp_bg_composite_forward" not in names_t
    for a, b in zip(outs_f, outs_t):
        assert rel_err(a, b.detach()) < 1e-5
    for a, b, p in zip(grads_f, grads_t, params):
        assert rel_err(a, b) < 2e-4, tuple(p.shape)
