"""Generate the golden fixtures under `tests/golden/` from the REFERENCE's own code.

Test infrastructure; runs only in the build container (needs `/root/reference`, which does not
exist on the GPU box -- the committed `.npz` files are what travels).

For every case it stores the inputs, random output cotangents, and the outputs + gradients of
  (a) the reference's naive PyTorch path (`lightplane_renderer_naive`, `lightplane_splatter_naive`,
      `lightplane_mlp_splatter_naive`), and
  (b) the reference's Triton kernels executed on CPU with TRITON_INTERPRET=1, from a scratch copy
      in /tmp in which `_floor(x) = x - x % 1` (grid_sample_util.py:12-14) is replaced by a true
      floor: under the installed Triton 3.6 float `%` lowers to C fmod, which truncates negative
      coordinates; the reference pins triton==2.1.0 where `%` was a floor-mod.
The reference sources are never copied into this repository.

Usage:  python oracle/make_golden.py            (writes tests/golden/*.npz)
"""

import os
import shutil
import sys

os.environ["TRITON_INTERPRET"] = "1"

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
SCRATCH = "/tmp/lp_ref_scratch"
OUT = os.path.join(REPO, "tests", "golden")


def _prepare_reference_copy():
    if os.path.exists(SCRATCH):
        shutil.rmtree(SCRATCH)
    shutil.copytree(os.path.join(REF, "lightplane"), os.path.join(SCRATCH, "lightplane"))
    p = os.path.join(SCRATCH, "lightplane", "triton_src", "shared", "grid_sample_util.py")
    src = open(p).read()
    assert "return x - x % 1" in src
    src = src.replace("return x - x % 1", "return tl.floor(x)")
    open(p, "w").write(src)
    sys.path.insert(0, SCRATCH)
    sys.path.insert(0, os.path.join(HERE, "_refshim"))


_prepare_reference_copy()

import numpy as np  # noqa: E402
import torch  # noqa: E402

import lightplane as ref  # noqa: E402  (the scratch copy of the reference)

assert ref.__file__.startswith(SCRATCH)


def make_rays(n, batch, enc_dim, seed):
    """Same distribution as the reference's test generator (tests/utils.py:230-268):
    rays start near the origin-centred cube, point roughly through it and cross its borders."""
    g = torch.Generator().manual_seed(seed)
    origins = torch.randn(n, 3, generator=g) / 3.0
    directions = -origins + 0.1 * torch.randn(n, 3, generator=g)
    near = torch.randn(n, generator=g) * 0.1 + 0.1
    far = torch.randn(n, generator=g).abs() * 0.1 + 3.0
    grid_idx = torch.randint(0, batch, (n,), generator=g)
    enc = torch.randn(n, enc_dim, generator=g) if enc_dim else None
    return directions, origins, grid_idx, near, far, enc


def grid_shapes(size, triplane):
    if not triplane:
        return [list(size)]
    out = []
    for i in range(3):
        s = list(size)
        s[i + 1] = 1
        out.append(s)
    return out


def run_renderer_case(name, *, n_rays, size, triplane, layers, hidden, color_chn, num_samples,
                      num_samples_inf=0, gain=1.0, mask_oob=False, contract=False, sigma=0.0,
                      noise_seed=0, scaffold_size=None, color_grid=False, seed=0, run_triton=True):
    torch.manual_seed(seed)
    B, C = size[0], size[4]
    n_t, n_o, n_c = layers
    dp = ref.init_decoder_params(
        device="cpu", n_layers_opacity=n_o, n_layers_trunk=n_t, n_layers_color=n_c, input_chn=C,
        hidden_chn=hidden, color_chn=color_chn, opacity_init_bias=-1.0,
        use_separate_color_grid=color_grid,
    )
    dp.mlp_params = (dp.mlp_params + 0.05 * torch.randn_like(dp.mlp_params)).requires_grad_(True)
    enc_dim = C if color_grid else hidden
    d, o, gi, nr, fr, enc = make_rays(n_rays, B, enc_dim, seed + 100)
    enc.requires_grad_(True)
    shapes = grid_shapes(size, triplane)
    grids = [torch.randn(s).requires_grad_(True) for s in shapes]
    cgrids = [torch.randn(s).requires_grad_(True) for s in shapes] if color_grid else None
    scaffold = None
    if scaffold_size is not None:
        scaffold = (torch.randn(B, *scaffold_size) > -0.3).float()
    cot = [torch.randn(n_rays), torch.randn(n_rays), torch.randn(n_rays, color_chn)]
    kwargs = dict(num_samples=num_samples, gain=gain, num_samples_inf=num_samples_inf,
                  mask_out_of_bounds_samples=mask_oob, contract_coords=contract,
                  inject_noise_sigma=sigma, inject_noise_seed=noise_seed, scaffold=scaffold,
                  color_grid=cgrids)

    def run(fn):
        rays = ref.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=enc)
        outs = fn(rays, grids, dp, **kwargs)
        loss = sum((c * v).sum() for c, v in zip(cot, outs))
        leaves = grids + [dp.mlp_params, enc] + (cgrids or [])
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        gg = torch.cat([g.reshape(-1, C) for g in grads[: len(grids)]], 0)
        res = dict(ray_length=outs[0], nlt=outs[1], features=outs[2], g_grid=gg,
                   g_mlp=grads[len(grids)], g_enc=grads[len(grids) + 1])
        if cgrids:
            res["g_color_grid"] = torch.cat([g.reshape(-1, C) for g in grads[len(grids) + 2:]], 0)
        return {k: v.detach().numpy() for k, v in res.items()}

    data = dict(
        directions=d, origins=o, grid_idx=gi.int(), near=nr, far=fr, encoding=enc.detach(),
        grid=torch.cat([g.detach().reshape(-1, C) for g in grids], 0),
        grid_sizes=np.array(shapes, dtype=np.int32), mlp_params=dp.mlp_params.detach(),
        n_hidden_trunk=dp.n_hidden_trunk, n_hidden_opacity=dp.n_hidden_opacity,
        n_hidden_color=dp.n_hidden_color, color_chn=np.int32(color_chn),
        cot_ray_length=cot[0], cot_nlt=cot[1], cot_features=cot[2],
        cfg=np.array([num_samples, num_samples_inf, int(mask_oob), int(contract), noise_seed], dtype=np.int64),
        cfg_f=np.array([gain, 1e-5, sigma], dtype=np.float64),
    )
    if cgrids:
        data["color_grid"] = torch.cat([g.detach().reshape(-1, C) for g in cgrids], 0)
    if scaffold is not None:
        data["scaffold"] = scaffold
    for k, v in run(ref.lightplane_renderer_naive).items():
        data["naive_" + k] = v
    if run_triton:
        for k, v in run(ref.lightplane_renderer).items():
            data["triton_" + k] = v
        for k in ("ray_length", "nlt", "features", "g_grid", "g_mlp", "g_enc"):
            a, b = data["naive_" + k], data["triton_" + k]
            err = np.abs(a - b).mean() / max(np.abs(a).mean(), 1e-12)
            print(f"  {name}: naive-vs-triton {k}: {err:.2e}")
            # Known reference-internal divergences: noise index bases differ when
            # N%16 != 0; with background samples the Triton kernels evaluate the disparity
            # schedule in fp32 ((d-1)*f+1 cancels catastrophically at f->1, ~1e-3 relative on the
            # last depths) while the naive path evaluates it in Python doubles.
            if not (sigma > 0 and n_rays % 16) and num_samples_inf == 0:
                assert err < 2e-4, (name, k, err)
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in data.items()})
    print("wrote", name)


def run_splatter_case(name, *, n_rays, size, triplane, feat_dim, num_samples, num_samples_inf=0,
                      mask_oob=False, contract=False, mlp=None, input_size=None, seed=0,
                      run_triton=True):
    torch.manual_seed(seed)
    B, C = size[0], size[4]
    d, o, gi, nr, fr, _ = make_rays(n_rays, B, 0, seed + 100)
    feat = torch.rand(n_rays, feat_dim).requires_grad_(True)
    shapes = grid_shapes(size, triplane)
    rows = sum(int(np.prod(s[:4])) for s in shapes)
    cot = torch.randn(rows, C)
    sp, in_grids, in_shapes = None, None, None
    if mlp is not None:
        n_layers, hidden = mlp
        sp = ref.init_splatter_params(device="cpu", n_layers=n_layers, input_chn=feat_dim,
                                      hidden_chn=hidden, out_chn=C)
        sp.mlp_params = (sp.mlp_params + 0.05 * torch.randn_like(sp.mlp_params)).requires_grad_(True)
        in_shapes = grid_shapes(input_size, triplane)
        in_grids = [torch.randn(s).requires_grad_(True) for s in in_shapes]
    kwargs = dict(num_samples=num_samples, num_samples_inf=num_samples_inf,
                  mask_out_of_bounds_samples=mask_oob, contract_coords=contract, return_list=False)

    def run(triton):
        rays = ref.Rays(directions=d, origins=o, grid_idx=gi, near=nr, far=fr, encoding=feat)
        if mlp is None:
            fn = ref.lightplane_splatter if triton else ref.lightplane_splatter_naive
            out = fn(rays, [tuple(s) for s in shapes], **kwargs)
            leaves = [feat]
        else:
            fn = ref.lightplane_mlp_splatter if triton else ref.lightplane_mlp_splatter_naive
            out = fn(rays, [tuple(s) for s in shapes], sp, in_grids, **kwargs)
            leaves = [feat, sp.mlp_params] + in_grids
        grads = torch.autograd.grad((out * cot).sum(), leaves)
        res = dict(out=out, g_feat=grads[0])
        if mlp is not None:
            res["g_mlp"] = grads[1]
            res["g_input_grid"] = torch.cat([g.reshape(-1, feat_dim) for g in grads[2:]], 0)
        return {k: v.detach().numpy() for k, v in res.items()}

    data = dict(directions=d, origins=o, grid_idx=gi.int(), near=nr, far=fr, feature=feat.detach(),
                out_sizes=np.array(shapes, dtype=np.int32), cot=cot,
                cfg=np.array([num_samples, num_samples_inf, int(mask_oob), int(contract)], dtype=np.int64))
    if mlp is not None:
        data.update(mlp_params=sp.mlp_params.detach(), n_hidden=sp.n_hidden,
                    input_grid=torch.cat([g.detach().reshape(-1, feat_dim) for g in in_grids], 0),
                    input_sizes=np.array(in_shapes, dtype=np.int32))
    for k, v in run(False).items():
        data["naive_" + k] = v
    if run_triton:
        for k, v in run(True).items():
            data["triton_" + k] = v
            a, b = data["naive_" + k], v
            err = np.abs(a - b).mean() / max(np.abs(a).mean(), 1e-12)
            print(f"  {name}: naive-vs-triton {k}: {err:.2e}")
            assert err < 2e-4, (name, k, err)
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in data.items()})
    print("wrote", name)


def run_module_case(name, *, n_side, size, triplane, hidden, num_samples, scaffold_res, seed=0, **mod_kw):
    """Fixtures from the reference's `LightplaneRenderer` MODULE (renderer_module.py), naive implementation on CPU:
    forward (harmonic ray embedding + Linear, background colour, alpha), `calculate_scaffold`,
    `eval_opacity_at_points`, `eval_decoder_at_points`, `get_decoder_params_list`."""
    torch.manual_seed(seed)
    B, C = size[0], size[4]
    m = ref.LightplaneRenderer(num_samples=num_samples, color_chn=3, grid_chn=C, mlp_hidden_chn=hidden,
                               opacity_init_bias=-1.0, use_naive_impl=True, **mod_kw)
    with torch.no_grad():
        m.mlp_params.add_(0.05 * torch.randn_like(m.mlp_params))
    shapes = grid_shapes(size, triplane)
    grids = [torch.randn(s) for s in shapes]
    n = n_side * n_side
    ys, xs = torch.meshgrid(torch.linspace(-0.5, 0.5, n_side), torch.linspace(-0.5, 0.5, n_side), indexing="ij")
    d = torch.stack([xs, ys, -torch.ones_like(xs)], -1).reshape(-1, 3)
    o = torch.tensor([0.1, -0.05, 2.2]).expand(n, 3).contiguous()
    near, far = torch.full((n,), 1.0), torch.full((n,), 3.4)
    gi = (torch.arange(n) * B // n).long()
    rays = ref.Rays(directions=d, origins=o, grid_idx=gi, near=near, far=far)
    length, alpha, feat = m(rays, grids)
    # threshold = the median opacity over the lattice, so that about half of the cells survive before the dilation
    lin = torch.linspace(-1, 1, scaffold_res)
    lat = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3)
    thr = float(m.eval_opacity_at_points(lat, torch.zeros(1, dtype=torch.long), grids).median())
    scaffold = m.calculate_scaffold(grids, [B, scaffold_res, scaffold_res, scaffold_res], "cpu", threshold=thr, dilate_scaffold=0)
    length_s, alpha_s, feat_s = m(rays, grids, scaffold=scaffold)
    g = torch.Generator().manual_seed(seed + 1)
    pts = torch.rand(12, 7, 3, generator=g) * 2.4 - 1.2
    pidx = torch.randint(0, B, (12,), generator=g)
    pdirs = torch.randn(12, 3, generator=g)
    opa = m.eval_opacity_at_points(pts, pidx, grids)
    opa2, col2 = m.eval_decoder_at_points(pts, pidx, None, grids, directions=pdirs)
    opa3, col3 = m.eval_decoder_at_points(pts, pidx, None, grids, scaffold=scaffold, directions=pdirs)
    plist = m.get_decoder_params_list()
    data = dict(
        directions=d, origins=o, grid_idx=gi.int(), near=near, far=far,
        grid=torch.cat([x.reshape(-1, C) for x in grids], 0), grid_sizes=np.array(shapes, dtype=np.int32),
        mlp_params=m.mlp_params.detach(), lin_w=m.harmonic_ray_embedding_linear.weight.detach(),
        lin_b=m.harmonic_ray_embedding_linear.bias.detach(), bg_color=m.bg_color,
        cfg=np.array([num_samples, hidden, scaffold_res, int(triplane)], dtype=np.int64), scaffold_threshold=np.float64(thr),
        out_length=length.detach(), out_alpha=alpha.detach(), out_features=feat.detach(),
        scaffold=scaffold, outs_length=length_s.detach(), outs_alpha=alpha_s.detach(), outs_features=feat_s.detach(),
        pts=pts, pts_idx=pidx.int(), pts_dirs=pdirs, pts_opacity=opa.detach(),
        dec_opacity=opa2.detach(), dec_features=col2.detach(), decs_opacity=opa3.detach(), decs_features=col3.detach(),
        plist_shapes=np.array([[len(grp)] + [int(np.prod(t.shape)) for t in grp] + [0] * (4 - len(grp)) for grp in plist], dtype=np.int64),
        plist_flat=torch.cat([t.reshape(-1) for grp in plist for t in grp]).detach(),
    )
    for k, v in mod_kw.items():
        data["kw_" + k] = np.array(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in data.items()})
    print("wrote", name, "scaffold occupancy", float(scaffold.mean()))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--modules-only" not in sys.argv:
        main_ops()
    run_module_case("module_triplane_bg", n_side=12, size=(2, 10, 10, 10, 16), triplane=True, hidden=32, num_samples=24,
                    scaffold_res=10, seed=21, bg_color=(0.1, 0.5, 0.9), gain=2.0, mask_out_of_bounds_samples=True)
    run_module_case("module_voxel_logT", n_side=10, size=(1, 8, 8, 8, 16), triplane=False, hidden=32, num_samples=16,
                    scaffold_res=8, seed=22, bg_color=0.25, return_log_transmittance=True, ray_embedding_num_harmonics=2)


def main_ops():
    V = (2, 6, 5, 4, 16)
    # renderer: voxel / triplane / every optional feature at least once
    run_renderer_case("render_voxel_222", n_rays=32, size=V, triplane=False, layers=(2, 2, 2),
                      hidden=16, color_chn=3, num_samples=8, seed=1)
    run_renderer_case("render_triplane_inf_gain", n_rays=32, size=V, triplane=True, layers=(2, 2, 2),
                      hidden=32, color_chn=3, num_samples=8, num_samples_inf=3, gain=3.0, seed=2)
    run_renderer_case("render_triplane_mask_layers", n_rays=48, size=V, triplane=True,
                      layers=(1, 3, 2), hidden=16, color_chn=4, num_samples=6, mask_oob=True, seed=3)
    run_renderer_case("render_voxel_contract_inf", n_rays=32, size=V, triplane=False,
                      layers=(3, 1, 1), hidden=16, color_chn=3, num_samples=6, num_samples_inf=4,
                      contract=True, seed=4)
    run_renderer_case("render_voxel_scaffold_pad35", n_rays=35, size=V, triplane=False,
                      layers=(2, 2, 2), hidden=16, color_chn=3, num_samples=8,
                      scaffold_size=(5, 4, 6), seed=5)
    run_renderer_case("render_colorgrid_noise", n_rays=32, size=V, triplane=True, layers=(0, 2, 2),
                      hidden=16, color_chn=3, num_samples=8, sigma=1.0, noise_seed=77,
                      color_grid=True, seed=6)
    run_renderer_case("render_c32_b1", n_rays=16, size=(1, 4, 5, 6, 32), triplane=True,
                      layers=(2, 2, 2), hidden=32, color_chn=3, num_samples=8, seed=7)
    # splatter
    run_splatter_case("splat_voxel", n_rays=32, size=V, triplane=False, feat_dim=16, num_samples=8,
                      num_samples_inf=3, seed=11)
    run_splatter_case("splat_triplane_mask_contract", n_rays=35, size=V, triplane=True, feat_dim=16,
                      num_samples=8, mask_oob=True, contract=True, seed=12)
    run_splatter_case("splat_mlp_voxel", n_rays=32, size=V, triplane=False, feat_dim=16,
                      num_samples=6, mlp=(2, 16), input_size=(2, 4, 5, 6, 16), seed=13)
    run_splatter_case("splat_mlp_triplane3", n_rays=16, size=(2, 6, 5, 4, 32), triplane=True,
                      feat_dim=16, num_samples=6, num_samples_inf=2, mlp=(3, 32),
                      input_size=(2, 4, 5, 6, 16), seed=14)


if __name__ == "__main__":
    main()
