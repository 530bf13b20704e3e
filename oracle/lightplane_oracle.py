"""ORACLE -- test infrastructure, NOT part of the product.

CPU restatement (plain PyTorch, autograd for gradients, fp32 or fp64) of the algorithm on the
Renderer / Splatter hot path of facebookresearch/lightplane.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; nothing under `lightplane_b200/` does.

Parity pinning: `oracle/make_golden.py` runs the reference's own code in the build container
(its naive PyTorch path, and its Triton kernels under TRITON_INTERPRET=1 with the `_floor` fix
described in SURVEY.md H2) and stores inputs + outputs + gradients under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against every stored vector.

Each function cites the reference lines it restates (paths relative to the reference repo).
Sampling is written as explicit corner gathers (the Triton formulation,
lightplane/triton_src/shared/grid_sample_util.py) rather than `F.grid_sample` (the naive
formulation, lightplane/naive_renderer.py:674-731) so that the two reference formulations and
this one are mutually independent.
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

# ------------------------------------------------------------------------------------------
# depth schedule, contraction, RNG
# ------------------------------------------------------------------------------------------


def ray_depths(near, far, num_samples: int, num_samples_inf: int, disparity_at_inf: float):
    """Depths `[N, S+S_inf]` and step lengths `delta` of the same shape.

    t_j = near + (far-near) * j/(S-1)                      (ray_util.py:54-58 depth_lin)
    background: t_k = far / ((d_inf - 1)(k+1)/S_inf + 1)   (ray_util.py:47-51 depth_inv_sphere)
    delta_0 = (far-near)/(S-1) (1 if S == 1), delta_j = t_j - t_{j-1}
    (renderer_fw.py:209-226; naive_renderer.py:218-219,239-257)."""
    lin = torch.linspace(0.0, 1.0, num_samples, dtype=near.dtype, device=near.device)
    depths = near[:, None] + lin[None, :] * (far - near)[:, None]
    if num_samples_inf > 0:
        # The naive reference evaluates 1/n_disp in Python doubles (naive_renderer.py:810-813) and
        # only then multiplies the fp32 `far`; in fp32 `(d-1)*f + 1` cancels catastrophically for
        # f -> 1 (the Triton kernels do that and are ~1e-3 off on the last depths).  The doubles
        # are the semantics kept here and in the CUDA kernels (host-side table of scales).
        scale = [
            1.0 / ((disparity_at_inf - 1.0) * ((k + 1) / num_samples_inf) + 1.0)
            for k in range(num_samples_inf)
        ]
        scale = torch.tensor(scale, dtype=near.dtype, device=near.device)
        depths = torch.cat([depths, far[:, None] * scale[None, :]], dim=1)
    if num_samples > 1:
        first = (far - near) / (num_samples - 1)
    else:
        first = torch.ones_like(near)
    delta = torch.cat([first[:, None], depths[:, 1:] - depths[:, :-1]], dim=1)
    return depths, delta


def contract_pi(x: torch.Tensor) -> torch.Tensor:
    """MERF contraction followed by x0.5 (ray_util.py:12-45; naive_renderer.py:796-807)."""
    n = x.abs().amax(dim=-1, keepdim=True)
    ax = x.abs()
    on_max = (ax - n).abs() <= 1e-8
    safe = torch.where(ax > 0, ax, torch.ones_like(ax))
    contracted = torch.where(on_max, (2.0 - 1.0 / safe) * (x / safe), x / torch.clamp(n, min=1e-30))
    return torch.where(n <= 1.0, x, contracted) * 0.5


_INT32_PRIME = 105097564


def _i32(x: torch.Tensor) -> torch.Tensor:
    """Wrap an int64 tensor to int32 two's-complement range."""
    return ((x + 2**31) % 2**32) - 2**31


def _hash32(x):
    """rand_util.py:38-44 `hash` on int32 with wrap-around multiply and arithmetic shifts."""
    for _ in range(2):
        x = _i32(((x >> 16) ^ x) * 0x45D9F3B)
    return (x >> 16) ^ x


def _pair_hash32(x, h):
    """rand_util.py:47-52 `pair_hash`: h ^= x; h = (h << 24) + h * 0x193 (int32 wrap)."""
    h = _i32(h ^ x)
    return _i32(_i32(h << 24) + _i32(h * 0x193))


def int_to_randn(x1: torch.Tensor, x2: torch.Tensor, seed: int) -> torch.Tensor:
    """Hash two int32 -> N(0,1) via Box-Muller, fp32 (rand_util.py:64-79, :108-145)."""
    x1, x2 = _i32(x1.long()), _i32(x2.long())
    s = _i32(torch.tensor(int(seed), dtype=torch.long))
    p = torch.tensor(_INT32_PRIME, dtype=torch.long)
    h1 = _pair_hash32(_pair_hash32(p, s), _hash32(x1))
    h2 = _pair_hash32(_pair_hash32(p, _i32(s + 1)), _hash32(x2))
    f32 = torch.float32
    denom = torch.tensor(4294967295.0 + 3.0, dtype=f32)
    u1 = ((h1.to(f32) + torch.tensor(2147483647.0, dtype=f32)) + 3.0) / denom
    u2 = ((h2.to(f32) + torch.tensor(2147483647.0, dtype=f32)) + 3.0) / denom
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(torch.tensor(6.28318530718, dtype=f32) * u2)


def sample_noise(num_rays: int, tot_samples: int, seed: int) -> torch.Tensor:
    """`[N, S_tot]` opacity noise.  Index convention of the Triton kernels
    (fwbw_util.py:66-70, renderer_fw.py:289-296): i1 = ray*S_tot + step + 1,
    i2 = i1 + N_padded*S_tot with N padded to a multiple of 16."""
    n_pad = ((num_rays + 15) // 16) * 16
    i1 = tot_samples * torch.arange(num_rays)[:, None] + torch.arange(tot_samples)[None, :] + 1
    i2 = i1 + n_pad * tot_samples
    return int_to_randn(i1.reshape(-1), i2.reshape(-1), seed).reshape(num_rays, tot_samples)


# ------------------------------------------------------------------------------------------
# grid-list addressing
# ------------------------------------------------------------------------------------------


def _grid_kind(D: int, H: int, W: int) -> str:
    """voxel if all of D,H,W > 1; else XY if D == 1, XZ if H == 1, else YZ
    (grid_sample_util.py:1111-1173)."""
    if (D - 1) * (H - 1) * (W - 1) > 0:
        return "voxel"
    if D == 1:
        return "xy"
    if H == 1:
        return "xz"
    return "yz"


def _axis_setup(p, size: int):
    """Continuous index, floor, fraction for one axis (align_corners=False);
    singleton axes are pinned to index 0 (grid_sample_util.py:231-247)."""
    i = ((p + 1.0) * 0.5) * size - 0.5
    if size <= 1:
        i = torch.zeros_like(i)
    i0 = torch.floor(i)
    return i0, i - i0


def _corner_terms(pts, grid_idx, size):
    """Yield (row_index [N,S] long, weight [N,S]) for the 8 (voxel) / 4 (plane) taps of one
    grid; out-of-range taps get weight 0 and a clamped index (= zero padding)
    (grid_sample_util.py:638-714, :780-1085)."""
    B, D, H, W, _ = size
    kind = _grid_kind(D, H, W)
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]
    b = grid_idx[:, None].long()
    if kind == "voxel":
        axes = [(x, W), (y, H), (z, D)]
    elif kind == "xy":
        axes = [(x, W), (y, H)]
    elif kind == "xz":
        axes = [(x, W), (z, D)]
    else:
        axes = [(y, H), (z, D)]
    setups = [_axis_setup(p, n) + (n,) for p, n in axes]
    terms = []
    for corner in range(2 ** len(axes)):
        w = torch.ones_like(x)
        idxs = []
        for a, (i0, frac, n) in enumerate(setups):
            hi = (corner >> a) & 1
            ia = i0 + hi
            w = w * (frac if hi else (1.0 - frac))
            w = w * ((ia >= 0) & (ia < n)).to(w.dtype)
            idxs.append(ia.clamp(0, n - 1).long())
        if kind == "voxel":
            ix, iy, iz = idxs
            row = ((b * D + iz) * H + iy) * W + ix
        elif kind == "xy":
            ix, iy = idxs
            row = (b * H + iy) * W + ix
        elif kind == "xz":
            ix, iz = idxs
            row = (b * D + iz) * W + ix
        else:
            iy, iz = idxs
            row = (b * D + iz) * H + iy
        terms.append((row, w))
    return terms


def _in_bounds(pts):
    return ((pts.abs() <= 1.0).all(dim=-1)).to(pts.dtype)


def _split_flat(flat: torch.Tensor, sizes: Sequence[Sequence[int]]):
    rows = [s[0] * s[1] * s[2] * s[3] for s in sizes]
    return flat.split(rows, dim=0)


def sample_grid_list(flat, sizes, grid_idx, pts, mask_oob: bool):
    """Sum over the grid-list of tri/bi-linear samples, `[N,S,C]`
    (grid_sample_util.py:1088-1216; naive_renderer.py:625-731)."""
    out = 0
    for g, size in zip(_split_flat(flat, sizes), sizes):
        for row, w in _corner_terms(pts, grid_idx, size):
            out = out + g[row] * w[..., None]
    if mask_oob:
        out = out * _in_bounds(pts)[..., None]
    return out


def sample_nearest(values, size, grid_idx, pts):
    """Nearest-neighbour lookup of a 1-channel voxel grid with OOB masking, `[N,S]`
    (grid_sample_util.py:717-777, round = floor(x + 0.5); naive_renderer.py:484-499)."""
    B, D, H, W = size[:4]
    b = grid_idx[:, None].long()
    idx, ok = [], torch.ones_like(pts[..., 0], dtype=torch.bool)
    for p, n in ((pts[..., 0], W), (pts[..., 1], H), (pts[..., 2], D)):
        i = torch.floor(((p + 1.0) * 0.5) * n - 0.5 + 0.5)
        if n <= 1:
            i = torch.zeros_like(i)
        ok = ok & (i >= 0) & (i < n)
        idx.append(i.clamp(0, n - 1).long())
    ix, iy, iz = idx
    row = ((b * D + iz) * H + iy) * W + ix
    return values.reshape(-1)[row] * ok.to(values.dtype) * _in_bounds(pts)


def splat_grid_list(flat_out, sizes, grid_idx, pts, feat, mask_oob: bool):
    """Adjoint of `sample_grid_list`: returns `flat_out + splat(feat)` where feat is `[N,S,C]`
    (grid_sample_util.py:40-206,1219-1246; naive_splatter.py:315-668)."""
    if mask_oob:
        feat = feat * _in_bounds(pts)[..., None]
    C = feat.shape[-1]
    outs, pos = [], 0
    for size in sizes:
        n_rows = size[0] * size[1] * size[2] * size[3]
        g = flat_out[pos : pos + n_rows]
        for row, w in _corner_terms(pts, grid_idx, size):
            g = g.index_add(0, row.reshape(-1), (feat * w[..., None]).reshape(-1, C))
        outs.append(g)
        pos += n_rows
    return torch.cat(outs, dim=0)


# ------------------------------------------------------------------------------------------
# MLPs
# ------------------------------------------------------------------------------------------


def split_mlp(flat: torch.Tensor, dims: Sequence[int]):
    """Weights-then-biases layout of one MLP (mlp_utils.py:691-721)."""
    ws, bs, pos = [], [], 0
    for i, o in zip(dims[:-1], dims[1:]):
        ws.append(flat[pos : pos + i * o].reshape(i, o))
        pos += i * o
    for o in dims[1:]:
        bs.append(flat[pos : pos + o])
        pos += o
    assert pos == flat.numel()
    return ws, bs


def mlp_numel(dims: Sequence[int]) -> int:
    return sum(i * o for i, o in zip(dims[:-1], dims[1:])) + sum(dims[1:])


def eval_mlp(x, ws, bs, relu_last: bool = False):
    """y = x@W+b with ReLU between layers (naive_renderer.py:758-776); the trunk additionally
    applies ReLU after its last layer (renderer_mlp_util.py:109-110; naive_renderer.py:399)."""
    for l, (w, b) in enumerate(zip(ws, bs)):
        x = x @ w + b
        if l < len(ws) - 1 or relu_last:
            x = torch.relu(x)
    return x


# ------------------------------------------------------------------------------------------
# Renderer
# ------------------------------------------------------------------------------------------


def render(
    directions, origins, grid_idx, near, far, encoding,
    grid_flat, grid_sizes, mlp_params,
    dims_trunk: Sequence[int], dims_opacity: Sequence[int], dims_color: Sequence[int],
    num_samples: int, gain: float, num_samples_inf: int = 0,
    mask_out_of_bounds_samples: bool = False, contract_coords: bool = False,
    disparity_at_inf: float = 1e-5, inject_noise_sigma: float = 0.0, inject_noise_seed: int = 0,
    scaffold: Optional[torch.Tensor] = None,
    color_grid_flat: Optional[torch.Tensor] = None, color_grid_sizes=None,
):
    """Emission-absorption render of a ray batch; returns
    `(ray_length [N], neg_log_transmittance [N], features [N, dims_color[-1]])`.

    Restates renderer_fw.py:85-375 / naive_renderer.py:216-325 (march + composite) and
    naive_renderer.py:328-501 (decoder).  Differentiable w.r.t. grid_flat, mlp_params, encoding,
    color_grid_flat through autograd."""
    depths, delta = ray_depths(near, far, num_samples, num_samples_inf, disparity_at_inf)
    pts = origins[:, None, :] + depths[..., None] * directions[:, None, :]
    if contract_coords:
        pts = contract_pi(pts)

    n_t, n_o = mlp_numel(dims_trunk) if len(dims_trunk) else 0, mlp_numel(dims_opacity)
    w_t, b_t = split_mlp(mlp_params[:n_t], dims_trunk) if n_t else ([], [])
    w_o, b_o = split_mlp(mlp_params[n_t : n_t + n_o], dims_opacity)
    w_c, b_c = split_mlp(mlp_params[n_t + n_o :], dims_color)

    sampled = sample_grid_list(grid_flat, grid_sizes, grid_idx, pts, mask_out_of_bounds_samples)
    if color_grid_flat is None:
        trunk = eval_mlp(sampled, w_t, b_t, relu_last=True)
        color_in = trunk
    else:  # relu-field mode (renderer_fw.py:267-316)
        trunk = torch.relu(sampled)
        color_in = torch.relu(
            sample_grid_list(color_grid_flat, color_grid_sizes, grid_idx, pts, mask_out_of_bounds_samples)
        )
    opacity_raw = eval_mlp(trunk, w_o, b_o)[..., 0]
    if inject_noise_sigma > 0.0:
        noise = sample_noise(directions.shape[0], depths.shape[1], inject_noise_seed)
        opacity_raw = opacity_raw + inject_noise_sigma * noise.to(opacity_raw.dtype)
    opacity = torch.nn.functional.softplus(opacity_raw)
    color = torch.sigmoid(eval_mlp(color_in + encoding[:, None, :], w_c, b_c))
    if scaffold is not None:
        occ = sample_nearest(scaffold, list(scaffold.shape) + [1], grid_idx, pts)
        opacity = opacity * occ
        color = color * occ[..., None]

    # compositing (renderer_fw.py:345-363; naive_renderer.py:303-316)
    nlt = torch.cumsum(delta * gain * opacity, dim=1)
    transmittance = torch.exp(-nlt)
    prev = torch.cat([torch.ones_like(transmittance[:, :1]), transmittance[:, :-1]], dim=1)
    weights = prev - transmittance
    ray_length = (weights * depths).sum(dim=1)
    features = (weights[..., None] * color).sum(dim=1)
    return ray_length, nlt[:, -1], features


# ------------------------------------------------------------------------------------------
# Splatter
# ------------------------------------------------------------------------------------------


def splat(
    directions, origins, grid_idx, near, far, feature,
    out_sizes, num_samples: int, num_samples_inf: int = 0,
    mask_out_of_bounds_samples: bool = False, contract_coords: bool = False,
    disparity_at_inf: float = 1e-5,
    mlp_params: Optional[torch.Tensor] = None, mlp_dims: Optional[Sequence[int]] = None,
    input_grid_flat: Optional[torch.Tensor] = None, input_sizes=None,
):
    """Splat per-ray features (optionally `MLP(sample(input_grid) + feature)`) at every
    ray-march sample; returns the flat normalised grid `[sum BDHW, C]`.

    Restates lightplane_splatter.py:343-584 + splatter_fw.py:71-309 /
    naive_splatter.py:185-289: out = F / clamp(Wt, 1e-5), Wt detached (it does not depend on
    any differentiable input)."""
    depths, _ = ray_depths(near, far, num_samples, num_samples_inf, disparity_at_inf)
    pts = origins[:, None, :] + depths[..., None] * directions[:, None, :]
    if contract_coords:
        pts = contract_pi(pts)
    S = depths.shape[1]
    if mlp_params is not None:
        ws, bs = split_mlp(mlp_params, mlp_dims)
        sampled = sample_grid_list(input_grid_flat, input_sizes, grid_idx, pts, mask_out_of_bounds_samples)
        feat = eval_mlp(sampled + feature[:, None, :], ws, bs)
    else:
        feat = feature[:, None, :].expand(-1, S, -1)
    C = out_sizes[0][4]
    rows = sum(s[0] * s[1] * s[2] * s[3] for s in out_sizes)
    acc = torch.zeros(rows, C, dtype=feat.dtype)
    acc = splat_grid_list(acc, out_sizes, grid_idx, pts, feat, mask_out_of_bounds_samples)
    ones = torch.ones(feat.shape[0], S, 1, dtype=feat.dtype)
    wsizes = [list(s[:4]) + [1] for s in out_sizes]
    wacc = splat_grid_list(torch.zeros(rows, 1, dtype=feat.dtype), wsizes, grid_idx, pts, ones,
                           mask_out_of_bounds_samples)
    return acc / torch.clamp(wacc.detach(), min=1e-5)
