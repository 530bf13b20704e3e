"""`LightplaneRenderer` -- the `torch.nn.Module` front-end of the renderer hot path.

Host-side mirror of the reference's `lightplane/renderer_module.py:38-667`: same constructor
arguments, same per-call overrides, same return values `(ray_length, alpha, features)`, same
`ValueError`s for ray-encoding misuse (:604-667).  The module owns `mlp_params`
(`nn.Parameter`) and the harmonic ray-embedding `Linear`, and calls the CUDA-bound
`lightplane_renderer`.
"""

from __future__ import annotations

import copy
import logging
from typing import Optional, Tuple

import torch

from .lightplane_renderer import lightplane_renderer
from .misc_utils import if_not_none_else, process_and_flatten_grid
from .module_ops import bg_composite, ray_embedding_linear, ray_embedding_supported
from .mlp_utils import DecoderParams, flattened_decoder_params_to_list, init_decoder_params
from .ray_utils import Rays, calc_harmonic_embedding, calc_harmonic_embedding_dim, jitter_near_far

logger = logging.getLogger(__name__)

_NO_NAIVE = (
    "use_naive_impl=True is not available in lightplane_b200: the PyTorch restatement of the"
    " algorithm lives in `oracle/` as test infrastructure and is never reachable from the"
    " product path. Use the reference package if you need its naive implementation."
)


class LightplaneRenderer(torch.nn.Module):
    """Renders `Rays` from a feature grid-list with learnable trunk / opacity / colour MLPs.

    Args mirror the reference (renderer_module.py:39-146): `num_samples`, `color_chn`,
    `grid_chn`, `mlp_hidden_chn`, layer counts, `use_separate_color_grid`, `opacity_init_bias`,
    `gain`, `bg_color`, `enable_direction_dependent_colors`, `ray_embedding_num_harmonics`,
    `num_samples_inf`, `mask_out_of_bounds_samples`, `contract_coords`, `disparity_at_inf`,
    `inject_noise_sigma`, `inject_noise_seed`, `rays_jitter_near_far`,
    `return_log_transmittance`; `triton_block_size` / `triton_num_warps` are accepted and ignored.
    """

    def __init__(
        self,
        num_samples: int,
        color_chn: int,
        grid_chn: int,
        mlp_hidden_chn: int,
        mlp_n_layers_opacity: int = 2,
        mlp_n_layers_trunk: int = 2,
        mlp_n_layers_color: int = 2,
        use_separate_color_grid: bool = False,
        opacity_init_bias: float = -5.0,
        gain: float = 1.0,
        bg_color=0.0,
        enable_direction_dependent_colors: bool = True,
        ray_embedding_num_harmonics: Optional[int] = 3,
        num_samples_inf: int = 0,
        mask_out_of_bounds_samples: bool = False,
        contract_coords: bool = False,
        disparity_at_inf: float = 1e-5,
        inject_noise_sigma: float = 0.0,
        inject_noise_seed: Optional[int] = None,
        rays_jitter_near_far: bool = False,
        return_log_transmittance: bool = False,
        triton_block_size: int = 16,
        triton_num_warps: int = 4,
        use_naive_impl: bool = False,
    ) -> None:
        super().__init__()
        if use_naive_impl:
            raise NotImplementedError(_NO_NAIVE)
        self.num_samples = num_samples
        self.color_chn = color_chn
        self.gain = gain
        self.opacity_init_bias = opacity_init_bias
        self.num_samples_inf = num_samples_inf
        self.mask_out_of_bounds_samples = mask_out_of_bounds_samples
        self.contract_coords = contract_coords
        self.disparity_at_inf = disparity_at_inf
        self.inject_noise_sigma = inject_noise_sigma
        self.inject_noise_seed = inject_noise_seed
        self.rays_jitter_near_far = rays_jitter_near_far
        self.return_log_transmittance = return_log_transmittance
        self.enable_direction_dependent_colors = enable_direction_dependent_colors
        self.ray_embedding_num_harmonics = ray_embedding_num_harmonics
        self.triton_block_size = triton_block_size
        self.triton_num_warps = triton_num_warps
        self.use_naive_impl = False

        if use_separate_color_grid and mlp_n_layers_trunk > 0:
            logger.warning(
                "Auto-setting mlp_n_layers_trunk=0 because a separate feature grid"
                " for colors is used (use_separate_color_grid=True)."
            )
            mlp_n_layers_trunk = 0

        params = init_decoder_params(
            device="cpu",
            n_layers_opacity=mlp_n_layers_opacity,
            n_layers_trunk=mlp_n_layers_trunk,
            n_layers_color=mlp_n_layers_color,
            input_chn=grid_chn,
            hidden_chn=mlp_hidden_chn,
            color_chn=color_chn,
            opacity_init_bias=opacity_init_bias,
            pad_color_channels_to_min_block_size=True,
            use_separate_color_grid=use_separate_color_grid,
        )
        self.mlp_params = torch.nn.Parameter(params.mlp_params)
        # Layer dims stay on the host (plain CPU tensors, not buffers): the launch path needs
        # them as python ints and must not read them back from the GPU on every call.
        self._n_hidden = (params.n_hidden_trunk, params.n_hidden_opacity, params.n_hidden_color)
        self.rays_encoding_dim = int(params.n_hidden_color[0])

        if ray_embedding_num_harmonics is not None:
            if not enable_direction_dependent_colors:
                raise ValueError(
                    "LightplaneRenderer's viewpoint dependent colors are disabled,"
                    " (enable_direction_dependent_colors=False), but `ray_embedding_num_harmonics`"
                    " is set. Set LightplaneRender.ray_embedding_num_harmonics = None if you"
                    " intended to disable viewpoint dependent colors."
                )
            self.harmonic_ray_embedding_linear = torch.nn.Linear(
                calc_harmonic_embedding_dim(ray_embedding_num_harmonics), self.rays_encoding_dim
            )
            torch.nn.init.xavier_uniform_(self.harmonic_ray_embedding_linear.weight)
            torch.nn.init.zeros_(self.harmonic_ray_embedding_linear.bias)

        self.register_buffer("bg_color", self._process_bg_color(bg_color))

    # reference-compatible views of the layer dims
    @property
    def n_hidden_trunk(self):
        return self._n_hidden[0]

    @property
    def n_hidden_opacity(self):
        return self._n_hidden[1]

    @property
    def n_hidden_color(self):
        return self._n_hidden[2]

    def get_decoder_params(self) -> DecoderParams:
        """Current decoder parameters as a `DecoderParams` (renderer_module.py:183-196)."""
        return DecoderParams(self.mlp_params, *self._n_hidden, self.color_chn)

    def _process_bg_color(self, bg_color) -> torch.Tensor:
        """Scalar or per-channel background colour -> `[color_chn]` tensor (:565-576)."""
        if isinstance(bg_color, (tuple, list)):
            bg = torch.tensor(bg_color, dtype=torch.float32)
        elif torch.is_tensor(bg_color):
            bg = bg_color.float()
        else:
            bg = torch.full((self.color_chn,), float(bg_color), dtype=torch.float32)
        if bg.numel() == 1:
            bg = bg.reshape(1).expand(self.color_chn).clone()
        assert bg.numel() == self.color_chn, "bg_color must have color_chn entries"
        return bg.reshape(self.color_chn)

    # ------------------------------------------------------------------------------------
    def forward(
        self,
        rays: Rays,
        feature_grid,
        color_feature_grid=None,
        scaffold: Optional[torch.Tensor] = None,
        grid_sizes=None,
        color_grid_sizes=None,
        bg_color=None,
        num_samples: Optional[int] = None,
        gain: Optional[float] = None,
        num_samples_inf: Optional[int] = None,
        mask_out_of_bounds_samples: Optional[bool] = None,
        contract_coords: Optional[bool] = None,
        disparity_at_inf: Optional[float] = None,
        inject_noise_sigma: Optional[float] = None,
        inject_noise_seed: Optional[int] = None,
        rays_jitter_near_far: Optional[bool] = None,
        return_log_transmittance: Optional[bool] = None,
        regenerate_code: Optional[bool] = None,
        ray_image_width: Optional[int] = None,
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Render; every keyword overrides the module default for this call
        (renderer_module.py:419-563).  Returns `(ray_length, alpha, features)` where alpha is
        `1 - exp(-NLT)` or `-NLT` when `return_log_transmittance`."""
        device = rays.device
        num_samples = if_not_none_else(num_samples, self.num_samples)
        gain = if_not_none_else(gain, self.gain)
        num_samples_inf = if_not_none_else(num_samples_inf, self.num_samples_inf)
        mask_oob = if_not_none_else(mask_out_of_bounds_samples, self.mask_out_of_bounds_samples)
        contract_coords = if_not_none_else(contract_coords, self.contract_coords)
        disparity_at_inf = if_not_none_else(disparity_at_inf, self.disparity_at_inf)
        sigma = if_not_none_else(inject_noise_sigma, self.inject_noise_sigma)
        seed = if_not_none_else(inject_noise_seed, self.inject_noise_seed)
        jitter = if_not_none_else(rays_jitter_near_far, self.rays_jitter_near_far)
        return_log_t = if_not_none_else(return_log_transmittance, self.return_log_transmittance)
        bg = self.bg_color if bg_color is None else self._process_bg_color(bg_color)
        bg = bg.to(device)

        _check_renderer_ray_encoding_input(
            rays.encoding,
            self.ray_embedding_num_harmonics,
            self.rays_encoding_dim,
            self.enable_direction_dependent_colors,
        )
        rays_enc = copy.copy(rays)  # shallow: fields are re-pointed, tensors shared
        rays_enc.encoding = self._get_ray_encoding(rays.encoding, rays.directions)
        if jitter:
            rays_enc.near, rays_enc.far = jitter_near_far(rays_enc.near, rays_enc.far, num_samples)

        ray_length, nlt, features = lightplane_renderer(
            rays_enc,
            feature_grid,
            self.get_decoder_params(),
            num_samples=num_samples,
            gain=gain,
            num_samples_inf=num_samples_inf,
            mask_out_of_bounds_samples=mask_oob,
            contract_coords=contract_coords,
            disparity_at_inf=disparity_at_inf,
            inject_noise_sigma=sigma,
            inject_noise_seed=seed,
            scaffold=scaffold,
            color_grid=color_feature_grid,
            grid_sizes=grid_sizes,
            color_grid_sizes=color_grid_sizes,
            ray_image_width=ray_image_width,
        )
        if features.is_cuda and not bg.requires_grad:
            alpha, features = bg_composite(nlt, features, bg, return_log_t)  # one launch each way (csrc/lp_ray_embed.cuh)
        else:
            transmittance = torch.exp(-nlt)
            features = features + transmittance[..., None] * bg
            alpha = -nlt if return_log_t else 1.0 - transmittance
        return ray_length, alpha, features

    # ------------------------------------------------------------------------------------
    def _get_ray_encoding(self, ray_encoding, directions) -> torch.Tensor:
        if ray_encoding is not None:
            assert not self.enable_direction_dependent_colors
            assert self.ray_embedding_num_harmonics is None
            return ray_encoding
        return self._get_ray_embedding(directions)

    def _get_ray_embedding(self, ray_directions: torch.Tensor) -> torch.Tensor:
        """Zero encoding, or Linear(harmonic(normalised direction)) (renderer_module.py:578-601)."""
        if not self.enable_direction_dependent_colors:
            return ray_directions.new_zeros(ray_directions.shape[0], self.rays_encoding_dim)
        assert self.ray_embedding_num_harmonics is not None
        lin = self.harmonic_ray_embedding_linear
        if (ray_directions.is_cuda and not ray_directions.requires_grad and ray_directions.dim() == 2
                and ray_embedding_supported(self.ray_embedding_num_harmonics, lin.out_features)):
            # fused normalize + embedding + Linear, forward and backward (csrc/lp_ray_embed.cuh)
            return ray_embedding_linear(ray_directions, lin.weight, lin.bias, self.ray_embedding_num_harmonics)
        unit = torch.nn.functional.normalize(ray_directions, dim=-1)
        return self.harmonic_ray_embedding_linear(
            calc_harmonic_embedding(unit, self.ray_embedding_num_harmonics)
        )

    # ------------------------------------------------------------------------------------
    def get_decoder_params_list(self):
        """`(weights_trunk, biases_trunk, weights_opacity, biases_opacity, weights_color, biases_color)`:
        the weight matrices / bias vectors inside `mlp_params` (renderer_module.py:255-284)."""
        return flattened_decoder_params_to_list(self.mlp_params, *self._n_hidden)

    def _points_as_rays(self, pts, pts_to_grid_idx, encoding):
        """Evaluation points as degenerate rays for the ray-march kernels: zero direction, origin = the point, so every
        sample of the "ray" is the point itself.  `pts` is `[n_rays, n_pts, 3]` with `pts_to_grid_idx [n_rays]` (the
        reference's layout) or flat `[P, 3]` with `[P]`."""
        if pts.ndim == 3:
            n_rays, n_pts, dim = pts.shape
            assert dim == 3 and tuple(pts_to_grid_idx.shape) == (n_rays,)
            flat = pts.reshape(-1, 3)
            idx = pts_to_grid_idx.repeat_interleave(n_pts)
            if encoding is not None:
                encoding = encoding.repeat_interleave(n_pts, dim=0)
            shape = (n_rays, n_pts)
        else:
            assert pts.ndim == 2 and pts.shape[1] == 3 and tuple(pts_to_grid_idx.shape) == (pts.shape[0],)
            flat, idx, shape = pts, pts_to_grid_idx, (pts.shape[0],)
        n = flat.shape[0]
        if encoding is None:
            encoding = flat.new_zeros(n, self.rays_encoding_dim)
        rays = Rays(directions=torch.zeros_like(flat), origins=flat.contiguous(), grid_idx=idx.to(torch.int32),
                    near=flat.new_zeros(n), far=flat.new_ones(n), encoding=encoding)
        return rays, shape

    def eval_opacity_at_points(
        self,
        pts: torch.Tensor,  # [n_rays, n_pts, 3]
        pts_to_grid_idx: torch.Tensor,  # [n_rays]
        feature_grid,
        scaffold: Optional[torch.Tensor] = None,
        gain: Optional[float] = None,
        mask_out_of_bounds_samples: Optional[bool] = None,
        grid_sizes=None,
        contract_coords: Optional[bool] = None,
    ) -> torch.Tensor:
        """`gain * softplus(opacity_raw)` (times the scaffold's occupancy) at 3-D points, `[n_rays, n_pts]`
        (renderer_module.py:302-346; flat `[P,3]` / `[P]` inputs are accepted too and give `[P]`).
        The reference evaluates this with its naive PyTorch decoder; here it goes THROUGH the ray-march kernel: each
        point is a zero-direction ray with near=0, far=1 and two samples (delta = 1 for both), whose negative log
        transmittance is `2 * gain * opacity`."""
        color_grid = feature_grid if self.n_hidden_trunk.numel() == 0 else None  # colour-grid mode: colours are irrelevant here
        rays, shape = self._points_as_rays(pts, pts_to_grid_idx, None)
        _, nlt, _ = lightplane_renderer(
            rays,
            feature_grid,
            self.get_decoder_params(),
            num_samples=2,
            gain=if_not_none_else(gain, self.gain),
            mask_out_of_bounds_samples=if_not_none_else(
                mask_out_of_bounds_samples, self.mask_out_of_bounds_samples
            ),
            contract_coords=bool(contract_coords) if contract_coords is not None else False,
            scaffold=scaffold,
            color_grid=color_grid,
            grid_sizes=grid_sizes,
            color_grid_sizes=grid_sizes if color_grid is not None else None,
        )
        return (0.5 * nlt).reshape(shape)

    def eval_decoder_at_points(
        self,
        pts: torch.Tensor,  # [n_rays, n_pts, 3]
        pts_to_grid_idx: torch.Tensor,  # [n_rays]
        rays_encoding: Optional[torch.Tensor],  # [n_rays, rays_encoding_dim]
        feature_grid,
        color_feature_grid=None,
        scaffold: Optional[torch.Tensor] = None,
        gain: Optional[float] = None,
        mask_out_of_bounds_samples: Optional[bool] = None,
        contract_coords: Optional[bool] = None,
        directions: Optional[torch.Tensor] = None,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Decoder outputs at 3-D points (renderer_module.py:183-253): `opacity [n_rays, n_pts]` = gain * softplus(raw)
        and `features [n_rays, n_pts, chn]` = sigmoid(colour logits), both times the scaffold's occupancy; `chn` is the
        width of the colour layer in the parameter layout (zero-padded to 16 at initialisation: those channels read
        sigmoid(0) = 0.5, as in the reference).  Evaluated THROUGH the ray-march kernels on degenerate rays (see `eval_opacity_at_points`):
        the opacity from a two-sample march; the colours from a one-sample march with a saturating gain of 1e30 (render weight
        1 - exp(-gain * softplus(raw)) == 1 in fp32 for raw > -60, and gain * softplus stays finite, so an unoccupied
        scaffold cell still multiplies it to an exact 0), which makes the rendered feature the decoder's colour."""
        n_rays = pts.shape[0]
        assert pts.ndim == 3 and pts.shape[2] == 3 and tuple(pts_to_grid_idx.shape) == (n_rays,)
        if rays_encoding is not None:
            assert tuple(rays_encoding.shape) == (n_rays, self.rays_encoding_dim)
        else:
            assert directions is not None, "Must pass one of (rays_encoding, directions)"
            assert tuple(directions.shape) == (n_rays, 3)
        enc = self._get_ray_encoding(rays_encoding, directions)
        mask_oob = if_not_none_else(mask_out_of_bounds_samples, self.mask_out_of_bounds_samples)
        contract = if_not_none_else(contract_coords, self.contract_coords)
        rays, shape = self._points_as_rays(pts, pts_to_grid_idx, enc)
        common = dict(mask_out_of_bounds_samples=mask_oob, contract_coords=contract, scaffold=scaffold,
                      color_grid=color_feature_grid)
        _, nlt, _ = lightplane_renderer(rays, feature_grid, self.get_decoder_params(), num_samples=2,
                                        gain=if_not_none_else(gain, self.gain), **common)
        # every channel of the colour layer as laid out in `mlp_params` (the padded ones too: whatever their weights hold)
        chn = int(self.n_hidden_color[-1])
        all_chn = DecoderParams(self.mlp_params, *self._n_hidden, chn)
        _, _, col = lightplane_renderer(rays, feature_grid, all_chn, num_samples=1, gain=1.0e30, **common)
        opacity = (0.5 * nlt).reshape(shape)
        return opacity, col.reshape(*shape, chn)

    @torch.no_grad()
    def calculate_scaffold(
        self,
        feature_grid,
        scaffold_size,  # [B, D, H, W]
        device,
        threshold: float = 1e-7,
        grid_sizes=None,
        dilate_scaffold: int = 2,
    ) -> torch.Tensor:
        """Occupancy scaffold `[B,D,H,W]` (1 = keep): opacity evaluated at voxel positions
        `linspace(-1, 1, size)` per axis, max-dilated, thresholded (renderer_module.py:348-417)."""
        B, D, H, W = (int(v) for v in scaffold_size)
        zs = torch.linspace(-1.0, 1.0, D, device=device)
        ys = torch.linspace(-1.0, 1.0, H, device=device)
        xs = torch.linspace(-1.0, 1.0, W, device=device)
        zz, yy, xx = torch.meshgrid(zs, ys, xs, indexing="ij")
        pts = torch.stack([xx, yy, zz], dim=-1).reshape(-1, 3)
        occ = torch.empty(B, D, H, W, device=device)
        for b in range(B):
            idx = torch.full((pts.shape[0],), b, device=device, dtype=torch.int32)
            occ[b] = self.eval_opacity_at_points(
                pts, idx, feature_grid, scaffold=None, gain=self.gain,
                mask_out_of_bounds_samples=self.mask_out_of_bounds_samples, grid_sizes=grid_sizes,
            ).reshape(D, H, W)
        if dilate_scaffold > 0:
            k = 2 * dilate_scaffold + 1
            occ = torch.nn.functional.max_pool3d(
                occ[:, None], kernel_size=k, padding=dilate_scaffold, stride=1
            )[:, 0]
        return (occ > threshold).float()


def _check_renderer_ray_encoding_input(
    ray_encoding, ray_embedding_num_harmonics, ray_encoding_dim: int,
    enable_direction_dependent_colors: bool,
) -> None:
    """Consistency of `rays.encoding` with the module configuration; raises `ValueError`
    in exactly the situations the reference does (renderer_module.py:604-667)."""
    if ray_encoding is not None and ray_encoding.shape[1] != ray_encoding_dim:
        raise ValueError(
            f"Ray encoding has a wrong dimension. Expected: {ray_encoding_dim}, got: {ray_encoding.shape[1]}"
        )
    if not enable_direction_dependent_colors:
        if ray_encoding is not None:
            raise ValueError(
                "LightplaneRenderer's viewpoint dependent colors are disabled"
                " (enable_direction_dependent_colors=False), but the `encoding` field of `rays` is"
                " set. Set rays.encoding=None if you intended to disable viewpoint dependent colors."
            )
        if ray_embedding_num_harmonics is not None:
            raise ValueError(
                "LightplaneRenderer's viewpoint dependent colors are disabled"
                " (enable_direction_dependent_colors=False), but `ray_embedding_num_harmonics` is"
                " set. Set it to None if you intended to disable viewpoint dependent colors."
            )
        return
    has_h, has_e = ray_embedding_num_harmonics is not None, ray_encoding is not None
    if has_h != has_e:
        return
    if not has_e:
        msg = (
            "rays.encoding is unset (=None), but the Lightplane module is not configured to"
            " compute harmonic ray embeddings (ray_embedding_num_harmonics is None)."
        )
    else:
        msg = (
            "rays.encoding is set, but the Lightplane module is configured to also compute"
            " harmonic ray embeddings (ray_embedding_num_harmonics is set)."
        )
    raise ValueError(
        msg
        + " Either let the module compute the embeddings (set ray_embedding_num_harmonics, leave"
        " rays.encoding=None) or supply your own [n_rays, ray_encoding_dim] rays.encoding and"
        " set ray_embedding_num_harmonics=None."
    )

