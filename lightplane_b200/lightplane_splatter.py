"""Functional Lightplane Splatter / MLP-Splatter bound to the sm_100a CUDA library.

Replaces the reference's `lightplane/lightplane_splatter.py`: `lightplane_splatter` (:31-164)
and `lightplane_mlp_splatter` (:167-338) keep their signatures; `LightplaneSplatterFunction`
(:341-700) now calls `lp_splat_*` / `lp_mlp_splat_*` of `include/lightplane_b200.h`.

The reference runs the forward ray march twice (features, then unit weights,
lightplane_splatter.py:505,539); here both grids are accumulated in one pass and normalised
in place (`feature / clamp(weight, 1e-5)`, :541,584).  The backward treats the weight grid as a
constant, as the reference does (:608).
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from .misc_utils import check_grid, process_and_flatten_grid, unflatten_grid
from .mlp_utils import SplatterParams

_byref = _cabi.byref


def _sizes_list(output_grid_size) -> List[List[int]]:
    if torch.is_tensor(output_grid_size):
        output_grid_size = output_grid_size.tolist()
    return [[int(v) for v in s] for s in output_grid_size]


def lightplane_splatter(
    rays,
    output_grid_size: List[Tuple[int, int, int, int, int]],
    # ------ config keys ------
    num_samples: int,
    num_samples_inf: int = 0,
    mask_out_of_bounds_samples: bool = False,
    contract_coords: bool = False,
    disparity_at_inf: float = 1e-5,
    return_list: bool = True,
    regenerate_code: bool = False,
    triton_block_size: int = 16,
    triton_num_warps: int = 4,
    process_group=None,
):
    """Splat `rays.encoding` into a zero-initialised grid-list of shapes `output_grid_size`
    at every sample of the ray march and normalise by the accumulated interpolation weights
    (reference semantics: lightplane_splatter.py:45-129).  Returns a list of `[B,D,H,W,C]`
    grids, or the flat `[sum BDHW, C]` tensor when `return_list=False`.

    `process_group` (extension, not in the reference): when rays are sharded across ranks, pass
    the `torch.distributed` group; the un-normalised feature and weight grids are then
    SUM-all-reduced before the normalisation, so every rank returns the full result."""
    del regenerate_code, triton_block_size, triton_num_warps
    sizes = _sizes_list(output_grid_size)
    out = LightplaneSplatterFunction.apply(
        rays.encoding, None, None,
        sizes, None, None,
        rays.directions, rays.origins, rays.grid_idx, rays.near, rays.far,
        int(num_samples), int(num_samples_inf), bool(mask_out_of_bounds_samples),
        bool(contract_coords), float(disparity_at_inf), process_group,
    )
    return list(unflatten_grid(out, sizes)) if return_list else out


def lightplane_mlp_splatter(
    rays,
    output_grid_size: List[Tuple[int, int, int, int, int]],
    mlp_params: SplatterParams,
    input_grid,
    # ------ config keys ------
    num_samples: int,
    num_samples_inf: int = 0,
    mask_out_of_bounds_samples: bool = False,
    contract_coords: bool = False,
    disparity_at_inf: float = 1e-5,
    input_grid_sizes: Optional[List[List[int]]] = None,
    return_list: bool = True,
    regenerate_code: bool = False,
    triton_block_size: int = 16,
    triton_num_warps: int = 4,
    process_group=None,
):
    """`input_grid -> sample -> + rays.encoding -> MLP -> splat -> output grid`
    (reference semantics: lightplane_splatter.py:184-288)."""
    del regenerate_code, triton_block_size, triton_num_warps
    sizes = _sizes_list(output_grid_size)
    n_hidden = [int(v) for v in mlp_params.n_hidden.tolist()]
    assert len(n_hidden) > 1, "mlp depth has to be bigger than 1 when using input_grid"
    assert input_grid is not None, "input_grid cannot be None when mlp_params is not None"
    check_grid(input_grid, input_grid_sizes)
    input_flat, _, input_sizes, _ = process_and_flatten_grid(input_grid, None, input_grid_sizes, None)
    out = LightplaneSplatterFunction.apply(
        rays.encoding, mlp_params.mlp_params, input_flat,
        sizes, input_sizes, n_hidden,
        rays.directions, rays.origins, rays.grid_idx, rays.near, rays.far,
        int(num_samples), int(num_samples_inf), bool(mask_out_of_bounds_samples),
        bool(contract_coords), float(disparity_at_inf), process_group,
    )
    return list(unflatten_grid(out, sizes)) if return_list else out


class LightplaneSplatterFunction(torch.autograd.Function):
    """autograd binding of the splatting kernels.  Differentiable inputs: the splatted per-ray
    feature, and for the MLP variant `mlp_params` and the flat input grid
    (reference: lightplane_splatter.py:677-700)."""

    @staticmethod
    def forward(
        ctx,
        splatting_feature: torch.Tensor,  # [N, E]
        mlp_params: Optional[torch.Tensor],
        input_grid: Optional[torch.Tensor],  # flat
        out_sizes: Sequence[Sequence[int]],
        input_sizes: Optional[Sequence[Sequence[int]]],
        n_hidden: Optional[Sequence[int]],
        directions, origins, grid_idx, near, far,
        num_samples: int,
        num_samples_inf: int,
        mask_out_of_bounds_samples: bool,
        contract_coords: bool,
        disparity_at_inf: float,
        process_group=None,
    ):
        lib = _cabi.get_lib()
        device = directions.device
        if device.type != "cuda":
            raise _cabi.LightplaneB200Error(
                f"lightplane_splatter runs on CUDA tensors only (no CPU fallback); got {device}"
            )
        use_mlp = mlp_params is not None
        num_rays = int(directions.shape[0])
        chn_out = int(out_sizes[0][4])
        assert all(int(s[4]) == chn_out for s in out_sizes), (
            "All output grids should have the same feature dimensions."
        )
        assert splatting_feature is not None and splatting_feature.ndim == 2
        assert splatting_feature.shape[0] == num_rays
        chn_feat = int(splatting_feature.shape[1])
        if use_mlp:
            assert input_grid is not None and input_sizes is not None and n_hidden is not None
            assert len(input_sizes) == len(out_sizes)
            assert n_hidden[0] == chn_feat and n_hidden[-1] == chn_out
            assert int(input_grid.shape[-1]) == chn_feat
            hidden = n_hidden[1] if len(n_hidden) > 2 else n_hidden[-1]
            assert all(h == hidden for h in n_hidden[1:-1])
            spec = _cabi.MlpSpec(len(n_hidden) - 1, n_hidden[0], hidden, n_hidden[-1])
        else:
            assert chn_out == chn_feat, "num_grid_channels should be the same as num_splatting_channels"
            spec = None
        assert tuple(directions.shape) == (num_rays, 3) and tuple(origins.shape) == (num_rays, 3)

        feat_c = _cabi.f32c(splatting_feature)
        dirs_c, orig_c = _cabi.f32c(directions), _cabi.f32c(origins)
        near_c, far_c = _cabi.f32c(near), _cabi.f32c(far)
        gidx_c = grid_idx.to(torch.int32).contiguous()
        mlp_c = _cabi.f32c(mlp_params) if use_mlp else None
        in_c = _cabi.f32c(input_grid) if use_mlp else None

        rows = sum(int(s[0]) * int(s[1]) * int(s[2]) * int(s[3]) for s in out_sizes)
        feature_grid = torch.zeros(rows, chn_out, device=device, dtype=torch.float32)
        weight_grid = torch.zeros(rows, 1, device=device, dtype=torch.float32)

        cfg = _cabi.make_cfg(
            num_samples, num_samples_inf, 1.0, disparity_at_inf, mask_out_of_bounds_samples,
            contract_coords, 0.0, 0, num_rays,
        )
        rays_s = _cabi.make_rays(dirs_c, orig_c, gidx_c, near_c, far_c, feat_c)
        out_s = _cabi.make_grid_list(feature_grid, out_sizes)
        in_s = _cabi.make_grid_list(in_c, input_sizes) if use_mlp else None
        stream = _cabi.stream_ptr(device)
        with torch.cuda.device(device):
            if num_rays > 0:
                if use_mlp:
                    st = _cabi.call(lib, "lp_mlp_splat_forward",
                        stream, _byref(cfg), _byref(spec), _byref(rays_s), None, _byref(in_s),
                        mlp_c.data_ptr(), _byref(out_s), weight_grid.data_ptr(),
                    )
                    _cabi.check(lib, st, "lp_mlp_splat_forward")
                else:
                    st = _cabi.call(lib, "lp_splat_forward",
                        stream, _byref(cfg), _byref(rays_s), None, _byref(out_s),
                        weight_grid.data_ptr(),
                    )
                    _cabi.check(lib, st, "lp_splat_forward")
            if process_group is not None:
                # ray-sharded splatting: reduce the UN-normalised accumulators (distributed.py)
                from .distributed import all_reduce_sum_

                all_reduce_sum_([feature_grid, weight_grid], process_group)
            st = lib.lp_splat_normalize(
                stream, feature_grid.data_ptr(), weight_grid.data_ptr(), rows, chn_out
            )
            _cabi.check(lib, st, "lp_splat_normalize")

        ctx.save_for_backward(weight_grid, feat_c, mlp_c, in_c, dirs_c, orig_c, gidx_c, near_c, far_c)
        ctx.lp = (cfg, spec, out_sizes, input_sizes)
        return feature_grid

    @staticmethod
    def backward(ctx, grad_feature_grid):
        lib = _cabi.get_lib()
        weight_grid, feat, mlp_params, input_grid, dirs, orig, gidx, near, far = ctx.saved_tensors
        cfg, spec, out_sizes, input_sizes = ctx.lp
        device = dirs.device
        num_rays = int(dirs.shape[0])
        use_mlp = spec is not None

        # weight grid is a constant w.r.t. every differentiable input (lightplane_splatter.py:608)
        g = (_cabi.f32c(grad_feature_grid) / weight_grid).contiguous()
        grad_feat = torch.empty_like(feat)
        grad_mlp = torch.zeros_like(mlp_params) if use_mlp else None
        grad_in = torch.zeros_like(input_grid) if use_mlp else None

        rays_s = _cabi.make_rays(dirs, orig, gidx, near, far, feat)
        g_s = _cabi.make_grid_list(g, out_sizes)
        stream = _cabi.stream_ptr(device)
        if num_rays > 0:
            with torch.cuda.device(device):
                if use_mlp:
                    in_s = _cabi.make_grid_list(input_grid, input_sizes)
                    st = _cabi.call(lib, "lp_mlp_splat_backward",
                        stream, _byref(cfg), _byref(spec), _byref(rays_s), None, _byref(in_s),
                        mlp_params.data_ptr(), _byref(g_s), grad_feat.data_ptr(),
                        grad_mlp.data_ptr(), grad_in.data_ptr(),
                    )
                    _cabi.check(lib, st, "lp_mlp_splat_backward")
                else:
                    st = _cabi.call(lib, "lp_splat_backward",
                        stream, _byref(cfg), _byref(rays_s), None, _byref(g_s), grad_feat.data_ptr()
                    )
                    _cabi.check(lib, st, "lp_splat_backward")
        else:
            grad_feat.zero_()
        return (grad_feat, grad_mlp, grad_in) + (None,) * 14
