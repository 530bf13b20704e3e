"""The renderer MODULE's glue as fused CUDA ops (kernels: csrc/lp_ray_embed.cuh).

* `ray_embedding_linear` = `Linear(calc_harmonic_embedding(F.normalize(directions)))`
  (reference: renderer_module.py:578-601, ray_utils.py:181-212) -- one launch each way;
* `bg_composite` = `features + exp(-nlt) * bg`, `alpha = 1 - exp(-nlt)` or `-nlt`
  (reference: renderer_module.py:552-563) -- one launch each way.

Both go through the C-ABI library (`_cabi.get_lib()` raises if it is missing: no fallback).
Neither differentiates w.r.t. ray directions or the background colour; `LightplaneRenderer` uses the
PyTorch composition of the same formulas when one of those requires a gradient.
"""

from __future__ import annotations

import torch

from . import _cabi


class _RayEmbedLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, directions, weight, bias, n_harmonics):
        lib = _cabi.get_lib()
        d, w = _cabi.f32c(directions.detach()), _cabi.f32c(weight.detach())
        b = _cabi.f32c(bias.detach()) if bias is not None else None
        n, out_dim = d.shape[0], w.shape[0]
        if d.dim() != 2 or d.shape[1] != 3:
            raise ValueError(f"directions must be [n, 3], got {tuple(d.shape)}")
        if w.shape[1] != 3 + 6 * n_harmonics:
            raise ValueError(f"weight must be [out, {3 + 6 * n_harmonics}], got {tuple(w.shape)}")
        enc = torch.empty(n, out_dim, device=d.device, dtype=torch.float32)
        st = _cabi.call(lib, "lp_ray_embed_forward", _cabi.stream_ptr(d.device), n, d.data_ptr(), int(n_harmonics),
                        w.data_ptr(), b.data_ptr() if b is not None else None, out_dim, enc.data_ptr())
        _cabi.check(lib, st, "lp_ray_embed_forward")
        ctx.save_for_backward(d)
        ctx.n_harmonics, ctx.has_bias, ctx.w_shape = int(n_harmonics), bias is not None, tuple(w.shape)
        return enc

    @staticmethod
    def backward(ctx, g_enc):
        (d,) = ctx.saved_tensors
        lib = _cabi.get_lib()
        g = _cabi.f32c(g_enc)
        g_w = torch.zeros(ctx.w_shape, device=d.device, dtype=torch.float32)
        g_b = torch.zeros(ctx.w_shape[0], device=d.device, dtype=torch.float32) if ctx.has_bias else None
        st = _cabi.call(lib, "lp_ray_embed_backward", _cabi.stream_ptr(d.device), d.shape[0], d.data_ptr(), ctx.n_harmonics,
                        g.data_ptr(), ctx.w_shape[0], g_w.data_ptr(), g_b.data_ptr() if g_b is not None else None)
        _cabi.check(lib, st, "lp_ray_embed_backward")
        return None, g_w, g_b, None


def ray_embedding_supported(n_harmonics: int, out_dim: int) -> bool:
    """Shapes the fused kernels take (lp_cabi.cu lp_embed_check): <= 10 harmonics, `out_dim` a multiple of 4 and
    at most 1024 (embedding column, float4 output chunk) pairs."""
    return 0 <= n_harmonics <= 10 and out_dim >= 4 and out_dim % 4 == 0 and (4 + 6 * n_harmonics) * (out_dim // 4) <= 1024


def ray_embedding_linear(directions: torch.Tensor, weight: torch.Tensor, bias, n_harmonics: int) -> torch.Tensor:
    """`[n, 3]` directions -> `[n, out]` ray encoding; gradients to `weight` and `bias` only."""
    return _RayEmbedLinear.apply(directions, weight, bias, int(n_harmonics))


class _BgComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, nlt, features, bg, return_log_t):
        lib = _cabi.get_lib()
        l, f, b = _cabi.f32c(nlt.detach()), _cabi.f32c(features.detach()), _cabi.f32c(bg.detach())
        n, c = f.shape
        if b.numel() != c:
            raise ValueError(f"bg_color has {b.numel()} channels, features {c}")
        alpha, out = torch.empty_like(l), torch.empty_like(f)
        st = _cabi.call(lib, "lp_bg_composite_forward", _cabi.stream_ptr(f.device), n, c, l.data_ptr(), f.data_ptr(),
                        b.data_ptr(), int(bool(return_log_t)), alpha.data_ptr(), out.data_ptr())
        _cabi.check(lib, st, "lp_bg_composite_forward")
        ctx.save_for_backward(l, b)
        ctx.log_t, ctx.c = int(bool(return_log_t)), c
        return alpha, out

    @staticmethod
    def backward(ctx, g_alpha, g_out):
        l, b = ctx.saved_tensors
        lib = _cabi.get_lib()
        ga = _cabi.f32c(g_alpha) if g_alpha is not None else None
        go = _cabi.f32c(g_out) if g_out is not None else None
        g_nlt = torch.empty_like(l)
        st = _cabi.call(lib, "lp_bg_composite_backward", _cabi.stream_ptr(l.device), l.shape[0], ctx.c, l.data_ptr(),
                        b.data_ptr(), ctx.log_t, ga.data_ptr() if ga is not None else None,
                        go.data_ptr() if go is not None else None, g_nlt.data_ptr())
        _cabi.check(lib, st, "lp_bg_composite_backward")
        return g_nlt, go, None, None


def bg_composite(nlt: torch.Tensor, features: torch.Tensor, bg: torch.Tensor, return_log_transmittance: bool):
    """`(alpha, features + exp(-nlt) * bg)`; gradients to `nlt` and `features` only."""
    return _BgComposite.apply(nlt, features, bg, bool(return_log_transmittance))
