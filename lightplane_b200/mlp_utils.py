"""MLP parameter containers and the flat parameter layout read by the CUDA kernels.

Host-side mirror of the reference's `lightplane/mlp_utils.py` (DecoderParams :20-128,
SplatterParams :131-185, init_decoder_params :188-295, init_splatter_params :298-339,
get_triton_function_input_dims :342-382, flatten_* :390-486, flattened_*_to_list :489-610).

Flat layout (the contract with `csrc/`): for each MLP in the order trunk, opacity, colour:
all weight matrices `W_l [in, out]` row-major (`y = x @ W_l + b_l`), then all bias vectors.
The colour head's last layer is zero-padded to `MIN_BLOCK_SIZE` = 16 output channels, exactly
as the reference does for its Triton kernels, so parameter tensors are interchangeable.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch

# Lower bound on channel counts inherited from the reference's Triton tiles
# (reference: lightplane/triton_src/shared/const.py:14-15).  Re-homed here because the
# Triton package is gone.
MIN_BLOCK_SIZE = 16


@dataclass
class DecoderParams:
    """Flat parameters of the renderer's decoder (trunk / opacity / colour MLPs).

    n_hidden_* are int32 tensors `[dim_in, dim_1, ..., dim_out]` per MLP (empty for a missing
    trunk); `color_chn` is the number of *useful* colour channels (the flat layout may carry
    more, zero-padded).  (reference: mlp_utils.py:20-128)
    """

    mlp_params: torch.Tensor
    n_hidden_trunk: torch.Tensor
    n_hidden_opacity: torch.Tensor
    n_hidden_color: torch.Tensor
    color_chn: int


@dataclass
class SplatterParams:
    """Flat parameters of the MLP splatter: weights then biases (mlp_utils.py:131-185)."""

    mlp_params: torch.Tensor
    n_hidden: torch.Tensor


# ----------------------------------------------------------------------------------------
# initialisation
# ----------------------------------------------------------------------------------------
def _xavier_mlp(n_layers, dim_in, dim_hidden, dim_out, device, last_bias=0.0):
    """Xavier-uniform weights with ReLU gain, zero biases except `last_bias` on the output
    layer (reference: mlp_utils.py:757-815)."""
    gain = torch.nn.init.calculate_gain("relu")
    ws, bs = [], []
    for l in range(n_layers):
        k = dim_in if l == 0 else dim_hidden
        n = dim_out if l == n_layers - 1 else dim_hidden
        w = torch.empty(k, n, device=device)
        torch.nn.init.xavier_uniform_(w, gain=gain)
        ws.append(w.contiguous())
        bs.append(torch.full((n,), last_bias if l == n_layers - 1 else 0.0, device=device))
    return ws, bs


def init_decoder_params(
    device,
    n_layers_opacity: int,
    n_layers_trunk: int,
    n_layers_color: int,
    input_chn: int = 32,
    hidden_chn: int = 32,
    color_chn: int = 3,
    opacity_init_bias: float = 0.0,
    pad_color_channels_to_min_block_size: bool = True,
    use_separate_color_grid: bool = False,
) -> DecoderParams:
    """Create Xavier-initialised decoder parameters (reference: mlp_utils.py:188-295)."""
    if n_layers_trunk > 0:
        assert not use_separate_color_grid, (
            "Cannot use trunk MLP with a separate color grid. Please set n_layers_trunk==0."
        )
        w_t, b_t = _xavier_mlp(n_layers_trunk, input_chn, hidden_chn, hidden_chn, device)
    else:
        w_t, b_t = [], []
    head_in = input_chn if use_separate_color_grid else hidden_chn
    w_o, b_o = _xavier_mlp(n_layers_opacity, head_in, hidden_chn, 1, device, opacity_init_bias)
    w_c, b_c = _xavier_mlp(n_layers_color, head_in, hidden_chn, color_chn, device)
    flat, nh_t, nh_o, nh_c = flatten_decoder_params(
        w_t, b_t, w_o, b_o, w_c, b_c, pad_color_channels_to_min_block_size
    )
    return DecoderParams(flat, nh_t, nh_o, nh_c, color_chn)


def init_splatter_params(
    device, n_layers: int, input_chn: int = 32, hidden_chn: int = 32, out_chn: int = 16
) -> SplatterParams:
    """Create Xavier-initialised MLP-splatter parameters (reference: mlp_utils.py:298-339)."""
    ws, bs = _xavier_mlp(n_layers, input_chn, hidden_chn, out_chn, device)
    flat, nh = flatten_splatter_params(ws, bs)
    return SplatterParams(flat, nh)


# ----------------------------------------------------------------------------------------
# flat <-> list conversions
# ----------------------------------------------------------------------------------------
def _layer_dims(weights: Sequence[torch.Tensor], device=None) -> torch.Tensor:
    """`[in, out_0, out_1, ...]` as int32 (empty tensor for no layers) (mlp_utils.py:724-740)."""
    if len(weights) == 0:
        return torch.zeros(0, dtype=torch.int32, device=device)
    dims = [int(weights[0].shape[0])] + [int(w.shape[1]) for w in weights]
    return torch.tensor(dims, dtype=torch.int32, device=device)


def _check_mlp_lists(ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor]) -> None:
    assert len(ws) == len(bs)
    prev_out = None
    for w, b in zip(ws, bs):
        assert w.ndim == 2 and b.ndim == 1 and w.device == b.device
        assert w.shape[1] == b.shape[0]
        if prev_out is not None:
            assert w.shape[0] == prev_out, "consecutive layers have inconsistent dims"
        prev_out = w.shape[1]


def _cat_flat(groups) -> torch.Tensor:
    parts = [t.reshape(-1) for grp in groups for t in grp]
    return torch.cat(parts, dim=0).contiguous()


def flatten_decoder_params(
    weights_trunk,
    biases_trunk,
    weights_opacity,
    biases_opacity,
    weights_color,
    biases_color,
    pad_color_channels_to_min_block_size: bool = True,
):
    """Concatenate the three MLPs into the flat layout; returns
    `(mlp_params, n_hidden_trunk, n_hidden_opacity, n_hidden_color)` (mlp_utils.py:390-456)."""
    weights_color, biases_color = list(weights_color), list(biases_color)
    if pad_color_channels_to_min_block_size:
        missing = MIN_BLOCK_SIZE - int(biases_color[-1].numel())
        if missing > 0:  # zero columns: padded channels never influence real ones
            weights_color[-1] = torch.nn.functional.pad(weights_color[-1], (0, missing))
            biases_color[-1] = torch.nn.functional.pad(biases_color[-1], (0, missing))
    for ws, bs in ((weights_trunk, biases_trunk), (weights_opacity, biases_opacity), (weights_color, biases_color)):
        _check_mlp_lists(ws, bs)
    flat = _cat_flat(
        [weights_trunk, biases_trunk, weights_opacity, biases_opacity, weights_color, biases_color]
    )
    assert flat.dtype == torch.float32
    dev = flat.device
    return (
        flat,
        _layer_dims(weights_trunk, dev),
        _layer_dims(weights_opacity, dev),
        _layer_dims(weights_color, dev),
    )


def flatten_splatter_params(weights, biases):
    """`(mlp_params, n_hidden)` for the single splatter MLP (mlp_utils.py:459-486)."""
    _check_mlp_lists(weights, biases)
    flat = _cat_flat([weights, biases])
    return flat, _layer_dims(weights, flat.device)


def _split_one_mlp(flat: torch.Tensor, n_hidden, transpose: bool = False):
    """Inverse of the per-MLP flattening (mlp_utils.py:691-721)."""
    dims = [int(v) for v in (n_hidden.tolist() if torch.is_tensor(n_hidden) else n_hidden)]
    if len(dims) == 0:
        assert flat.numel() == 0
        return [], []
    ins, outs = dims[:-1], dims[1:]
    n_w = sum(i * o for i, o in zip(ins, outs))
    assert flat.numel() == n_w + sum(outs), "mlp_params size does not match n_hidden"
    ws, pos = [], 0
    for i, o in zip(ins, outs):
        w = flat[pos : pos + i * o].reshape(i, o)
        ws.append(w.t().contiguous() if transpose else w)
        pos += i * o
    bs = []
    for o in outs:
        bs.append(flat[pos : pos + o])
        pos += o
    return ws, bs


def _mlp_numel(n_hidden) -> int:
    dims = [int(v) for v in (n_hidden.tolist() if torch.is_tensor(n_hidden) else n_hidden)]
    return sum(i * o for i, o in zip(dims[:-1], dims[1:])) + sum(dims[1:])


def flattened_decoder_params_to_list(
    mlp_params: torch.Tensor,
    n_hidden_trunk: torch.Tensor,
    n_hidden_opacity: torch.Tensor,
    n_hidden_color: torch.Tensor,
    transpose: bool = False,
):
    """Flat tensor -> `(weights_trunk, biases_trunk, weights_opacity, biases_opacity,
    weights_color, biases_color)`; inverse of `flatten_decoder_params` (mlp_utils.py:489-560)."""
    n_t, n_o = _mlp_numel(n_hidden_trunk), _mlp_numel(n_hidden_opacity)
    w_t, b_t = _split_one_mlp(mlp_params[:n_t], n_hidden_trunk, transpose)
    w_o, b_o = _split_one_mlp(mlp_params[n_t : n_t + n_o], n_hidden_opacity, transpose)
    w_c, b_c = _split_one_mlp(mlp_params[n_t + n_o :], n_hidden_color, transpose)
    return w_t, b_t, w_o, b_o, w_c, b_c


def flattened_triton_decoder_to_list(
    mlp_params: torch.Tensor,
    n_layers_trunk: int,
    n_layers_opacity: int,
    n_layers_color: int,
    input_chn: int,
    hidden_chn: int,
    color_chn: int,
):
    """Same as above but from layer counts / channel sizes (mlp_utils.py:563-610)."""

    def dims(d_in, d_out, n_layers):
        if n_layers <= 0:
            return torch.zeros(0, dtype=torch.int32, device=mlp_params.device)
        d = [d_in] + [hidden_chn] * (n_layers - 1) + [d_out]
        return torch.tensor(d, dtype=torch.int32, device=mlp_params.device)

    return flattened_decoder_params_to_list(
        mlp_params,
        dims(input_chn, hidden_chn, n_layers_trunk),
        dims(hidden_chn, 1, n_layers_opacity),
        dims(hidden_chn, color_chn, n_layers_color),
    )


def get_triton_function_input_dims(n_hidden_trunk, n_hidden_opacity, n_hidden_color):
    """`(hidden_trunk, hidden_opacity, hidden_color, layers_trunk, layers_opacity,
    layers_color, num_render_channels)` -- the name is kept from the reference
    (mlp_utils.py:342-382) because callers import it; nothing Triton is involved here.
    All hidden layers of one MLP must share one width."""

    def to_list(t):
        return [int(v) for v in (t.tolist() if torch.is_tensor(t) else t)]

    t, o, c = to_list(n_hidden_trunk), to_list(n_hidden_opacity), to_list(n_hidden_color)
    if len(t) == 0:
        hid_t, n_t = 0, 0
    else:
        hid_t, n_t = t[1], len(t) - 1
        assert all(v == hid_t for v in t[1:]), "all trunk layers must have the same width"
    hid_o, hid_c = o[1], c[1]
    assert all(v == hid_o for v in o[1:-1]), "all hidden opacity layers must have the same width"
    assert all(v == hid_c for v in c[1:-1]), "all hidden color layers must have the same width"
    return hid_t, hid_o, hid_c, n_t, len(o) - 1, len(c) - 1, c[-1]
