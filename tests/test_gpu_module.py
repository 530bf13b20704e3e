"""The nn.Module callers against fixtures produced by the REFERENCE's own modules (oracle/make_golden.py
`run_module_case`: `LightplaneRenderer` with use_naive_impl=True on CPU): forward incl. harmonic ray embedding,
background colour and alpha / log-transmittance; `calculate_scaffold`; `eval_opacity_at_points`;
`eval_decoder_at_points`; `get_decoder_params_list`.  And `LightplaneMLPSplatter.forward` against the oracle."""
import numpy as np
import pytest
import torch

from _golden import load_case, rel_err

CASES = ["module_triplane_bg", "module_voxel_logT"]


def _module(c, device):
    import lightplane_b200 as lp

    num_samples, hidden, _, _ = (int(v) for v in c["cfg"])
    kw = {k[3:]: (v.tolist() if v.ndim else v.item()) for k, v in c.items() if k.startswith("kw_")}
    if isinstance(kw.get("bg_color"), list):
        kw["bg_color"] = tuple(kw["bg_color"])
    C = int(c["grid_sizes"][0][4])
    m = lp.LightplaneRenderer(num_samples=num_samples, color_chn=3, grid_chn=C, mlp_hidden_chn=hidden,
                              opacity_init_bias=-1.0, **kw).to(device)
    with torch.no_grad():
        m.mlp_params.copy_(c["mlp_params"].to(device))
        m.harmonic_ray_embedding_linear.weight.copy_(c["lin_w"].to(device))
        m.harmonic_ray_embedding_linear.bias.copy_(c["lin_b"].to(device))
    return m


def _grids(c, device):
    rows = [int(np.prod(s[:4])) for s in c["grid_sizes"]]
    return [g.reshape([int(v) for v in s]).to(device) for g, s in zip(torch.split(c["grid"], rows), c["grid_sizes"])]


@pytest.mark.parametrize("name", CASES)
def test_module_decoder_params_list_matches_reference(name):
    """CPU: the parameter views of `get_decoder_params_list` are the reference's, tensor by tensor."""
    c = load_case(name)
    m = _module(c, "cpu")
    flat = torch.cat([t.reshape(-1) for grp in m.get_decoder_params_list() for t in grp])
    assert torch.equal(flat, c["plist_flat"])
    for grp, row in zip(m.get_decoder_params_list(), c["plist_shapes"]):
        assert len(grp) == int(row[0]) and [t.numel() for t in grp] == [int(v) for v in row[1:1 + len(grp)]]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_renderer_module_vs_reference_module(name):
    import lightplane_b200 as lp

    dev = "cuda"
    c = load_case(name)
    m, grids = _module(c, dev), _grids(c, dev)
    f = lambda k: c[k].to(dev)
    rays = lp.Rays(directions=f("directions"), origins=f("origins"), grid_idx=f("grid_idx"), near=f("near"), far=f("far"))
    length, alpha, feat = m(rays, grids)
    for got, key in ((length, "out_length"), (alpha, "out_alpha"), (feat, "out_features")):
        assert rel_err(got, c[key]) < 2e-4, (name, key, rel_err(got, c[key]))
    # scaffold: same occupancy grid (cells whose opacity sits within rounding of the threshold may flip)
    B = int(c["grid_sizes"][0][0])
    res = int(c["cfg"][2])
    scaf = m.calculate_scaffold(grids, [B, res, res, res], dev, threshold=float(c["scaffold_threshold"]), dilate_scaffold=0)
    assert scaf.shape == c["scaffold"].shape
    assert float((scaf.cpu() != c["scaffold"]).float().mean()) < 2e-3
    length, alpha, feat = m(rays, grids, scaffold=c["scaffold"].to(dev))
    for got, key in ((length, "outs_length"), (alpha, "outs_alpha"), (feat, "outs_features")):
        assert rel_err(got, c[key]) < 2e-4, (name, key, rel_err(got, c[key]))
    # point evaluations, the reference's [n_rays, n_pts, 3] layout
    pts, pidx, pdirs = f("pts"), torch.as_tensor(c["pts_idx"]).to(dev).long(), f("pts_dirs")
    opa = m.eval_opacity_at_points(pts, pidx, grids)
    assert opa.shape == c["pts_opacity"].shape and rel_err(opa, c["pts_opacity"]) < 2e-4
    opa2, col2 = m.eval_decoder_at_points(pts, pidx, None, grids, directions=pdirs)
    assert col2.shape == c["dec_features"].shape
    errs = (rel_err(opa2, c["dec_opacity"]), rel_err(col2[..., :3], c["dec_features"][..., :3]), rel_err(col2[..., 3:], c["dec_features"][..., 3:]))
    assert max(errs) < 2e-4, (errs, (col2[..., :3].cpu() - c["dec_features"][..., :3]).abs().amax(-1))
    opa3, col3 = m.eval_decoder_at_points(pts, pidx, None, grids, scaffold=c["scaffold"].to(dev), directions=pdirs)
    assert rel_err(opa3, c["decs_opacity"]) < 2e-4 and rel_err(col3, c["decs_features"]) < 2e-4


@pytest.mark.gpu
def test_mlp_splatter_module_forward_vs_oracle():
    """`LightplaneMLPSplatter.forward` (splatter_module.py:164-331) end to end against the oracle."""
    import lightplane_b200 as lp
    from oracle import lightplane_oracle as O

    dev = "cuda"
    torch.manual_seed(4)
    C_in, C_out, S = 16, 16, 24
    sizes = [(1, 10, 12, 14, C_out)]
    in_sizes = [(1, 6, 6, 6, C_in)]
    m = lp.LightplaneMLPSplatter(num_samples=S, grid_chn=C_out, input_grid_chn=C_in, mlp_hidden_chn=32, mlp_n_layers=2).to(dev)
    n = 400
    o = torch.randn(n, 3) / 3
    d = -o + 0.1 * torch.randn(n, 3)
    near, far = torch.full((n,), 0.1), torch.full((n,), 3.0)
    gi = torch.zeros(n, dtype=torch.int64)
    feat = torch.rand(n, C_in)
    in_grid = [torch.randn(in_sizes[0])]
    rays = lp.Rays(directions=d.to(dev), origins=o.to(dev), grid_idx=gi.to(dev), near=near.to(dev), far=far.to(dev),
                   encoding=feat.to(dev))
    out = m(rays, sizes, [g.to(dev) for g in in_grid])
    out = out if torch.is_tensor(out) else torch.cat([g.reshape(-1, C_out) for g in out], 0)
    f = lambda t: t.detach().double().cpu()
    want = O.splat(f(d), f(o), gi, f(near), f(far), f(feat), [list(s) for s in sizes], num_samples=S,
                   mlp_params=f(m.mlp_params), mlp_dims=[int(v) for v in m.n_hidden], input_grid_flat=f(in_grid[0]).reshape(-1, C_in),
                   input_sizes=[list(s) for s in in_sizes])
    assert rel_err(out.reshape(-1, C_out), want) < 2e-4, rel_err(out.reshape(-1, C_out), want)
