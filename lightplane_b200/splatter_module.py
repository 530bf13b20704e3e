"""`LightplaneSplatter` / `LightplaneMLPSplatter` -- module front-ends of the splatter path.

Host-side mirror of the reference's `lightplane/splatter_module.py` (LightplaneSplatter :25-161,
LightplaneMLPSplatter :164-331, `_check_splatter_ray_encoding_input` :334-348).
"""

from __future__ import annotations

import copy
from typing import List, Optional, Tuple

import torch

from .lightplane_splatter import lightplane_mlp_splatter, lightplane_splatter
from .misc_utils import if_not_none_else
from .mlp_utils import SplatterParams, init_splatter_params
from .ray_utils import Rays, jitter_near_far
from .renderer_module import _NO_NAIVE


class _SplatterBase(torch.nn.Module):
    def __init__(
        self, num_samples, num_samples_inf, mask_out_of_bounds_samples, contract_coords,
        disparity_at_inf, rays_jitter_near_far, triton_block_size, triton_num_warps,
        use_naive_impl,
    ):
        super().__init__()
        if use_naive_impl:
            raise NotImplementedError(_NO_NAIVE)
        self.num_samples = num_samples
        self.num_samples_inf = num_samples_inf
        self.mask_out_of_bounds_samples = mask_out_of_bounds_samples
        self.contract_coords = contract_coords
        self.disparity_at_inf = disparity_at_inf
        self.rays_jitter_near_far = rays_jitter_near_far
        self.triton_block_size = triton_block_size
        self.triton_num_warps = triton_num_warps
        self.use_naive_impl = False

    def _resolve(self, rays, num_samples, num_samples_inf, mask_oob, contract, disparity, jitter):
        cfg = dict(
            num_samples=if_not_none_else(num_samples, self.num_samples),
            num_samples_inf=if_not_none_else(num_samples_inf, self.num_samples_inf),
            mask_out_of_bounds_samples=if_not_none_else(mask_oob, self.mask_out_of_bounds_samples),
            contract_coords=if_not_none_else(contract, self.contract_coords),
            disparity_at_inf=if_not_none_else(disparity, self.disparity_at_inf),
        )
        _check_splatter_ray_encoding_input(rays.encoding, self.rays_encoding_dim)
        rays = copy.copy(rays)
        if if_not_none_else(jitter, self.rays_jitter_near_far):
            rays.near, rays.far = jitter_near_far(rays.near, rays.far, cfg["num_samples"])
        return rays, cfg


class LightplaneSplatter(_SplatterBase):
    """Splats `rays.encoding` into a zero-initialised grid-list (splatter_module.py:25-161)."""

    def __init__(
        self,
        num_samples: int,
        grid_chn: int,
        num_samples_inf: int = 0,
        mask_out_of_bounds_samples: bool = False,
        contract_coords: bool = False,
        disparity_at_inf: float = 1e-5,
        rays_jitter_near_far: bool = False,
        triton_block_size: int = 16,
        triton_num_warps: int = 4,
        use_naive_impl: bool = False,
    ):
        super().__init__(
            num_samples, num_samples_inf, mask_out_of_bounds_samples, contract_coords,
            disparity_at_inf, rays_jitter_near_far, triton_block_size, triton_num_warps,
            use_naive_impl,
        )
        self.rays_encoding_dim = grid_chn

    def get_splatter_params(self) -> Optional[SplatterParams]:
        return None

    def forward(
        self,
        rays: Rays,
        grid_size: List[Tuple[int, int, int, int, int]],
        num_samples: Optional[int] = None,
        num_samples_inf: Optional[int] = None,
        mask_out_of_bounds_samples: Optional[bool] = None,
        contract_coords: Optional[bool] = None,
        disparity_at_inf: Optional[float] = None,
        rays_jitter_near_far: Optional[bool] = None,
        return_list: bool = True,
        regenerate_code: bool = False,
    ):
        rays, cfg = self._resolve(
            rays, num_samples, num_samples_inf, mask_out_of_bounds_samples, contract_coords,
            disparity_at_inf, rays_jitter_near_far,
        )
        return lightplane_splatter(rays=rays, output_grid_size=grid_size, return_list=return_list, **cfg)


class LightplaneMLPSplatter(_SplatterBase):
    """Samples `input_grid`, adds `rays.encoding`, applies a learnable MLP and splats the
    result (splatter_module.py:164-331)."""

    def __init__(
        self,
        num_samples: int,
        grid_chn: int,
        input_grid_chn: int = 32,
        mlp_hidden_chn: int = 32,
        mlp_n_layers: int = 2,
        num_samples_inf: int = 0,
        mask_out_of_bounds_samples: bool = False,
        contract_coords: bool = False,
        disparity_at_inf: float = 1e-5,
        rays_jitter_near_far: bool = False,
        triton_block_size: int = 16,
        triton_num_warps: int = 4,
        use_naive_impl: bool = False,
    ):
        super().__init__(
            num_samples, num_samples_inf, mask_out_of_bounds_samples, contract_coords,
            disparity_at_inf, rays_jitter_near_far, triton_block_size, triton_num_warps,
            use_naive_impl,
        )
        assert input_grid_chn is not None, "input_grid_chn must be provided"
        p = init_splatter_params(
            device="cpu", n_layers=mlp_n_layers, input_chn=input_grid_chn,
            hidden_chn=mlp_hidden_chn, out_chn=grid_chn,
        )
        self.mlp_params = torch.nn.Parameter(p.mlp_params)
        self._n_hidden = p.n_hidden  # host tensor (see LightplaneRenderer)
        self.rays_encoding_dim = input_grid_chn

    @property
    def n_hidden(self):
        return self._n_hidden

    def get_splatter_params(self) -> SplatterParams:
        return SplatterParams(self.mlp_params, self._n_hidden)

    def forward(
        self,
        rays: Rays,
        grid_size: List[Tuple[int, int, int, int, int]],
        input_grid,
        num_samples: Optional[int] = None,
        num_samples_inf: Optional[int] = None,
        mask_out_of_bounds_samples: Optional[bool] = None,
        contract_coords: Optional[bool] = None,
        disparity_at_inf: Optional[float] = None,
        input_grid_sizes=None,
        rays_jitter_near_far: Optional[bool] = None,
        return_list: bool = True,
        regenerate_code: bool = False,
    ):
        assert input_grid is not None, "input_grid must be provided"
        rays, cfg = self._resolve(
            rays, num_samples, num_samples_inf, mask_out_of_bounds_samples, contract_coords,
            disparity_at_inf, rays_jitter_near_far,
        )
        return lightplane_mlp_splatter(
            rays=rays, output_grid_size=grid_size, mlp_params=self.get_splatter_params(),
            input_grid=input_grid, input_grid_sizes=input_grid_sizes, return_list=return_list, **cfg,
        )


def _check_splatter_ray_encoding_input(ray_encoding, ray_encoding_dim: int) -> None:
    """`ValueError` for a missing / mis-sized encoding (splatter_module.py:334-348)."""
    if ray_encoding is None:
        raise ValueError(
            "The encoding field of input rays is None. However, the Splatter requires an"
            " encoding for input rays."
        )
    if ray_encoding.shape[1] != ray_encoding_dim:
        raise ValueError(
            f"Ray encoding has a wrong dimension. Expected: {ray_encoding_dim}, got: {ray_encoding.shape[1]}"
        )
