"""Helpers shared by the parity tests: load a golden case, run the oracle on it, error norms."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names(prefix):
    return sorted(
        os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz"))
    )


def load_case(name, device="cpu", dtype=torch.float32):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind == "f" and v.ndim > 0 and not k.startswith("cfg"):
            out[k] = torch.from_numpy(v.astype(np.float32)).to(device=device, dtype=dtype)
        elif k in ("grid_idx",):
            out[k] = torch.from_numpy(v.astype(np.int32)).to(device)
        else:
            out[k] = v
    return out


def rel_err(a, b):
    """mean|a-b| / mean|b| -- the robust norm of SURVEY.md 8d."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().mean() / b.abs().mean().clamp_min(1e-30))


def max_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def renderer_cfg(c):
    num_samples, num_samples_inf, mask_oob, contract, noise_seed = (int(v) for v in c["cfg"])
    gain, disp, sigma = (float(v) for v in c["cfg_f"])
    return dict(num_samples=num_samples, num_samples_inf=num_samples_inf,
                mask_out_of_bounds_samples=bool(mask_oob), contract_coords=bool(contract),
                inject_noise_seed=noise_seed, gain=gain, disparity_at_inf=disp,
                inject_noise_sigma=sigma)


def oracle_render_case(c, dtype=torch.float64):
    """Run oracle.render on a loaded renderer case; returns outputs and grads (dict)."""
    from oracle import lightplane_oracle as O

    cfg = renderer_cfg(c)
    f = lambda t: t.detach().to(dtype=dtype, device="cpu")
    grid = f(c["grid"]).requires_grad_(True)
    mlp = f(c["mlp_params"]).requires_grad_(True)
    enc = f(c["encoding"]).requires_grad_(True)
    cgrid = f(c["color_grid"]).requires_grad_(True) if "color_grid" in c else None
    scaffold = f(c["scaffold"]) if "scaffold" in c else None
    sizes = [[int(v) for v in s] for s in c["grid_sizes"]]
    color_chn = int(c["color_chn"])
    outs = O.render(
        f(c["directions"]), f(c["origins"]), c["grid_idx"].cpu().long(), f(c["near"]), f(c["far"]),
        enc, grid, sizes, mlp,
        [int(v) for v in c["n_hidden_trunk"]], [int(v) for v in c["n_hidden_opacity"]],
        [int(v) for v in c["n_hidden_color"]],
        scaffold=scaffold, color_grid_flat=cgrid, color_grid_sizes=sizes if cgrid is not None else None,
        **cfg,
    )
    ray_length, nlt, feats = outs[0], outs[1], outs[2][:, :color_chn]
    loss = (f(c["cot_ray_length"]) * ray_length).sum() + (f(c["cot_nlt"]) * nlt).sum() + (
        f(c["cot_features"]) * feats).sum()
    leaves = [grid, mlp, enc] + ([cgrid] if cgrid is not None else [])
    grads = torch.autograd.grad(loss, leaves)
    res = dict(ray_length=ray_length, nlt=nlt, features=feats, g_grid=grads[0], g_mlp=grads[1],
               g_enc=grads[2])
    if cgrid is not None:
        res["g_color_grid"] = grads[3]
    return {k: v.detach() for k, v in res.items()}


def splat_cfg(c):
    num_samples, num_samples_inf, mask_oob, contract = (int(v) for v in c["cfg"])
    return dict(num_samples=num_samples, num_samples_inf=num_samples_inf,
                mask_out_of_bounds_samples=bool(mask_oob), contract_coords=bool(contract))


def oracle_splat_case(c, dtype=torch.float64):
    from oracle import lightplane_oracle as O

    f = lambda t: t.detach().to(dtype=dtype, device="cpu")
    feat = f(c["feature"]).requires_grad_(True)
    sizes = [[int(v) for v in s] for s in c["out_sizes"]]
    kw = splat_cfg(c)
    leaves = [feat]
    if "mlp_params" in c:
        mlp = f(c["mlp_params"]).requires_grad_(True)
        ing = f(c["input_grid"]).requires_grad_(True)
        kw.update(mlp_params=mlp, mlp_dims=[int(v) for v in c["n_hidden"]], input_grid_flat=ing,
                  input_sizes=[[int(v) for v in s] for s in c["input_sizes"]])
        leaves += [mlp, ing]
    out = O.splat(f(c["directions"]), f(c["origins"]), c["grid_idx"].cpu().long(), f(c["near"]),
                  f(c["far"]), feat, sizes, **kw)
    grads = torch.autograd.grad((out * f(c["cot"])).sum(), leaves)
    res = dict(out=out, g_feat=grads[0])
    if "mlp_params" in c:
        res.update(g_mlp=grads[1], g_input_grid=grads[2])
    return {k: v.detach() for k, v in res.items()}


def regrid_case(c, order=(0, 1, 2), sizes=None, seed=9):
    """Copy of a triplane renderer case with its grid LIST re-ordered (`order`) or given other plane sizes (`sizes`, fresh
    random features): the kernels' triplane fast path must not depend on the order, and must step aside for planes that
    do not belong to one volume."""
    c = dict(c)
    old = [[int(v) for v in q] for q in c["grid_sizes"]]
    C = old[0][4]
    if sizes is not None:
        g = torch.Generator().manual_seed(seed)
        new = [list(q) for q in sizes]
        c["grid"] = torch.randn(sum(q[0] * q[1] * q[2] * q[3] for q in new), C, generator=g)
    else:
        rows = [q[0] * q[1] * q[2] * q[3] for q in old]
        parts = list(torch.split(c["grid"], rows, 0))
        new = [old[i] for i in order]
        c["grid"] = torch.cat([parts[i] for i in order], 0)
    c["grid_sizes"] = np.array(new)
    return c


def plain_splat_case(n=70, channels=32, **kw):
    """A plain (no MLP) splatter case: `synthetic_splat_case` without its MLP entries."""
    c = synthetic_splat_case(n=n, c_in=channels, c_out=channels, **kw)
    for k in ("mlp_params", "n_hidden", "input_grid", "input_sizes"):
        c.pop(k)
    return c


def coherent_case(c, n=64, pixel=0.02, seed=5, batch_blocks=True, mask_oob=None, plane=None, scaffold_res=None,
                  origin=(0.1, -0.15, -2.2), near=1.0, far=3.4):
    """A copy of golden case `c` whose rays are replaced by `n` neighbouring pixels of a pinhole camera
    (the layout real renders have, and the one the backward kernel's warp-level scatter aggregation is
    built for); expected values then come from the oracle.  `plane`: optionally resize every grid to
    plane x plane texels with fresh random contents."""
    g = torch.Generator().manual_seed(seed)
    c = dict(c)
    w = int(n ** 0.5)
    ii = torch.arange(n)
    px, py = (ii % w).float() - w / 2, (ii // w).float() - w / 2
    d = torch.stack([px * pixel + 0.03, py * pixel - 0.02, torch.ones(n)], -1)
    c["directions"] = d / d.norm(dim=-1, keepdim=True)
    c["origins"] = torch.tensor(list(origin)).expand(n, 3).contiguous()
    c["near"] = torch.full((n,), float(near))
    c["far"] = torch.full((n,), float(far))
    B = int(c["grid_sizes"][0][0])
    c["grid_idx"] = ((ii * B) // n).int() if batch_blocks else torch.randint(0, B, (n,), generator=g).int()
    c["encoding"] = torch.randn(n, c["encoding"].shape[1], generator=g)
    c["cot_ray_length"] = torch.randn(n, generator=g)
    c["cot_nlt"] = torch.randn(n, generator=g)
    c["cot_features"] = torch.randn(n, c["cot_features"].shape[1], generator=g)
    if mask_oob is not None:
        cfg = c["cfg"].copy()
        cfg[2] = int(mask_oob)
        c["cfg"] = cfg
    if plane is not None:
        sizes = np.array([[int(v) if int(v) == 1 or j in (0, 4) else plane for j, v in enumerate(s)] for s in c["grid_sizes"]])
        c["grid_sizes"] = sizes
        rows = int(sum(int(np.prod(s[:4])) for s in sizes))
        c["grid"] = torch.randn(rows, int(sizes[0][4]), generator=g)
    if scaffold_res is not None:  # occupancy scaffold: a ball around the origin plus a few random cells, so that the
        r = scaffold_res            # first and last samples of every ray are in empty space (whole-group skips)
        ax = (torch.arange(r).float() + 0.5) / r * 2 - 1
        zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
        ball = ((xx ** 2 + yy ** 2 + zz ** 2) < 0.45).float()
        B = int(c["grid_sizes"][0][0])
        c["scaffold"] = torch.stack([torch.maximum(ball, (torch.rand(r, r, r, generator=g) > 0.97).float()) for _ in range(B)])
    return c


def synthetic_case(n=96, C=16, hidden=32, layers=(0, 2, 2), color_grid=True, plane=6, batch=2, samples=6, samples_inf=2,
                   mask_oob=1, contract=0, sigma=0.0, seed=3, pixel=0.03):
    """A renderer case built from scratch (decoder from `init_decoder_params`, camera-like rays) in the dict layout
    of the golden files, for configurations the golden set lacks; expected values come from the oracle."""
    import lightplane_b200 as lp

    g = torch.Generator().manual_seed(seed)
    nt, no, nc = layers
    torch.manual_seed(seed)
    dp = lp.init_decoder_params("cpu", no, nt, nc, input_chn=C, hidden_chn=hidden, color_chn=3, opacity_init_bias=-1.0,
                                use_separate_color_grid=color_grid)
    sizes = np.array([[batch, 1, plane, plane + 1, C], [batch, plane - 1, 1, plane + 1, C], [batch, plane - 1, plane, 1, C]])
    rows = int(sum(int(np.prod(s[:4])) for s in sizes))
    c = dict(
        grid=torch.randn(rows, C, generator=g), grid_sizes=sizes,
        mlp_params=dp.mlp_params.detach() + 0.05 * torch.randn(dp.mlp_params.shape, generator=g),
        n_hidden_trunk=dp.n_hidden_trunk.numpy(), n_hidden_opacity=dp.n_hidden_opacity.numpy(),
        n_hidden_color=dp.n_hidden_color.numpy(), color_chn=np.int32(3),
        cfg=np.array([samples, samples_inf, mask_oob, contract, 11]), cfg_f=np.array([1.5, 1e-5, sigma]),
        encoding=torch.zeros(n, C if color_grid else hidden), cot_features=torch.zeros(n, 3),
    )
    if color_grid:
        c["color_grid"] = torch.randn(rows, C, generator=g)
    return coherent_case(c, n=n, pixel=pixel, seed=seed)


def synthetic_splat_case(n=96, c_in=16, c_out=16, hidden=32, layers=2, res=6, batch=2, samples=6, samples_inf=2,
                         mask_oob=1, contract=0, seed=4, pixel=0.03, triplane=False):
    """An MLP-splatter case built from scratch in the dict layout of the golden files."""
    import lightplane_b200 as lp

    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    sp = lp.init_splatter_params("cpu", n_layers=layers, input_chn=c_in, hidden_chn=hidden, out_chn=c_out)

    def shapes(C):
        if triplane:
            return np.array([[batch, 1, res, res + 1, C], [batch, res - 1, 1, res + 1, C], [batch, res - 1, res, 1, C]])
        return np.array([[batch, res - 1, res, res + 1, C]])

    out_sizes, in_sizes = shapes(c_out), shapes(c_in)
    rows = lambda sz: int(sum(int(np.prod(q[:4])) for q in sz))
    w = int(n ** 0.5)
    ii = torch.arange(n)
    px, py = (ii % w).float() - w / 2, (ii // w).float() - w / 2
    d = torch.stack([px * pixel + 0.03, py * pixel - 0.02, torch.ones(n)], -1)
    return dict(
        directions=d / d.norm(dim=-1, keepdim=True), origins=torch.tensor([0.1, -0.15, -2.2]).expand(n, 3).contiguous(),
        near=torch.full((n,), 1.0), far=torch.full((n,), 3.4), grid_idx=((ii * batch) // n).int(),
        feature=torch.rand(n, c_in, generator=g), out_sizes=out_sizes, input_sizes=in_sizes,
        input_grid=torch.randn(rows(in_sizes), c_in, generator=g),
        mlp_params=sp.mlp_params.detach() + 0.05 * torch.randn(sp.mlp_params.shape, generator=g),
        n_hidden=sp.n_hidden.numpy(), cfg=np.array([samples, samples_inf, mask_oob, contract]),
        cot=torch.randn(rows(out_sizes), c_out, generator=g),
    )
