"""CPU tests of the host-side mirror of the reference interface (no GPU, no kernels)."""
import math

import pytest
import torch

import lightplane_b200 as lp
from lightplane_b200 import misc_utils, mlp_utils
from lightplane_b200.renderer_module import _check_renderer_ray_encoding_input


def _rays(n=10, enc=None):
    return lp.Rays(directions=torch.randn(n, 3), origins=torch.randn(n, 3), grid_idx=torch.zeros(n, dtype=torch.long),
                   near=torch.zeros(n), far=torch.ones(n), encoding=enc)


def test_rays_padding_indexing_and_validation():
    r = _rays(35, torch.randn(35, 8))
    p, extra = r.pad_to_block_size(16)
    assert extra == 13 and p.directions.shape == (48, 3) and p.encoding.shape == (48, 8)
    assert (p.directions[35:] == 0).all() and p.grid_idx.dtype == torch.long
    same, zero = p.pad_to_block_size(16)
    assert zero == 0 and same is p
    sub = r[5:9]
    assert sub.near.shape == (4,) and sub.encoding.shape == (4, 8)
    assert r.to("cpu") is r and r.clone().directions is not r.directions
    with pytest.raises(AssertionError):
        lp.Rays(directions=torch.randn(4, 3), origins=torch.randn(4, 3), grid_idx=torch.zeros(4),  # float idx
                near=torch.zeros(4), far=torch.ones(4))
    with pytest.raises(AssertionError):
        lp.Rays(directions=torch.randn(4, 3), origins=torch.randn(5, 3), grid_idx=torch.zeros(4, dtype=torch.long),
                near=torch.zeros(4), far=torch.ones(4))


def test_harmonic_embedding_layout():
    d = torch.nn.functional.normalize(torch.randn(7, 3), dim=-1)
    e = lp.calc_harmonic_embedding(d, 3)
    assert e.shape == (7, lp.calc_harmonic_embedding_dim(3)) == (7, 21)
    assert torch.allclose(e[:, -3:], d)
    # layout: [sin(d_x*2^k) k=0..2, sin(d_y..), sin(d_z..), cos(...) same order, d]
    assert torch.allclose(e[:, 0], torch.sin(d[:, 0]), atol=1e-6)
    assert torch.allclose(e[:, 4], torch.sin(2 * d[:, 1]), atol=1e-6)
    assert torch.allclose(e[:, 9 + 2], torch.cos(4 * d[:, 0]), atol=1e-5)
    assert lp.calc_harmonic_embedding(d, 0) is d
    n, f = lp.jitter_near_far(torch.zeros(100), torch.ones(100), 10)
    assert torch.allclose(f - n, torch.ones(100)) and n.abs().max() <= 0.1 + 1e-6


def test_decoder_params_layout_and_roundtrip():
    dp = lp.init_decoder_params("cpu", n_layers_opacity=2, n_layers_trunk=2, n_layers_color=2, input_chn=16,
                                hidden_chn=32, color_chn=3)
    assert dp.mlp_params.numel() == 4273  # 1600 + 1089 + 1584
    assert dp.n_hidden_trunk.tolist() == [16, 32, 32] and dp.n_hidden_color.tolist() == [32, 32, 16]
    assert dp.n_hidden_trunk.dtype == torch.int32 and dp.color_chn == 3
    wt, bt, wo, bo, wc, bc = lp.flattened_decoder_params_to_list(
        dp.mlp_params, dp.n_hidden_trunk, dp.n_hidden_opacity, dp.n_hidden_color)
    assert [tuple(w.shape) for w in wt] == [(16, 32), (32, 32)] and tuple(wo[-1].shape) == (32, 1)
    assert tuple(wc[-1].shape) == (32, 16) and (wc[-1][:, 3:] == 0).all() and (bc[-1][3:] == 0).all()
    flat2, *_ = lp.flatten_decoder_params(wt, bt, wo, bo, wc, bc, pad_color_channels_to_min_block_size=True)
    assert torch.equal(flat2, dp.mlp_params)
    assert lp.get_triton_function_input_dims(dp.n_hidden_trunk, dp.n_hidden_opacity, dp.n_hidden_color) == (
        32, 32, 32, 2, 2, 2, 16)
    # opacity bias initialisation and weights-then-biases order
    dp2 = lp.init_decoder_params("cpu", 1, 1, 1, input_chn=16, hidden_chn=16, color_chn=3, opacity_init_bias=-5.0)
    _, _, wo2, bo2, _, _ = lp.flattened_decoder_params_to_list(dp2.mlp_params, dp2.n_hidden_trunk,
                                                               dp2.n_hidden_opacity, dp2.n_hidden_color)
    assert float(bo2[0]) == -5.0 and tuple(wo2[0].shape) == (16, 1)
    with pytest.raises(AssertionError):
        lp.init_decoder_params("cpu", 2, 2, 2, use_separate_color_grid=True)
    cg = lp.init_decoder_params("cpu", 2, 0, 2, input_chn=16, hidden_chn=32, use_separate_color_grid=True)
    assert cg.n_hidden_trunk.numel() == 0 and cg.n_hidden_opacity.tolist() == [16, 32, 1]
    assert lp.get_triton_function_input_dims(cg.n_hidden_trunk, cg.n_hidden_opacity, cg.n_hidden_color)[3] == 0


def test_splatter_params():
    sp = lp.init_splatter_params("cpu", n_layers=3, input_chn=16, hidden_chn=64, out_chn=32)
    assert sp.n_hidden.tolist() == [16, 64, 64, 32]
    assert sp.mlp_params.numel() == 16 * 64 + 64 * 64 + 64 * 32 + 64 + 64 + 32
    ws, bs = mlp_utils._split_one_mlp(sp.mlp_params, sp.n_hidden)
    f2, nh = lp.flatten_splatter_params(ws, bs)
    assert torch.equal(f2, sp.mlp_params) and nh.tolist() == sp.n_hidden.tolist()


def test_grid_flatten_and_checks():
    grids = [torch.randn(2, 1, 5, 4, 16), torch.randn(2, 6, 1, 4, 16), torch.randn(2, 6, 5, 1, 16)]
    flat, sizes = lp.flatten_grid(grids)
    assert flat.shape == (2 * (20 + 24 + 30), 16) and sizes.dtype == torch.int32 and sizes.shape == (3, 5)
    back = lp.unflatten_grid(flat, sizes)
    assert all(torch.equal(a, b) for a, b in zip(grids, back))
    f, c, s, cs = misc_utils.process_and_flatten_grid(grids, None)
    assert torch.equal(f, flat) and c is None and s == [list(g.shape) for g in grids] and cs is None
    f2, _, s2, _ = misc_utils.process_and_flatten_grid(flat, None, s)
    assert f2.data_ptr() == flat.data_ptr() and s2 == s
    with pytest.raises(AssertionError):
        misc_utils.check_grid_and_color_grid(flat, None, None)  # flat grid without sizes
    with pytest.raises(AssertionError):
        misc_utils.check_grid_and_color_grid(grids, [g[:1] for g in grids])  # batch mismatch
    with pytest.raises(NotImplementedError):
        misc_utils.check_grid_and_color_grid("nope", None)
    assert misc_utils.is_in_bounds(torch.tensor([[0.5, -1.0, 1.0], [0.0, 1.01, 0.0]])).squeeze(-1).tolist() == [True, False]
    assert misc_utils.pad_feature_to_block_size(torch.ones(5, 3), 4).shape == (8, 3)


def test_renderer_module_construction_and_errors():
    m = lp.LightplaneRenderer(num_samples=8, color_chn=3, grid_chn=16, mlp_hidden_chn=32, bg_color=(0.0, 0.5, 1.0))
    assert isinstance(m.mlp_params, torch.nn.Parameter) and m.mlp_params.numel() == 4273
    assert m.rays_encoding_dim == 32 and m.harmonic_ray_embedding_linear.in_features == 21
    assert m.bg_color.tolist() == [0.0, 0.5, 1.0]
    dp = m.get_decoder_params()
    assert dp.color_chn == 3 and dp.n_hidden_trunk.device.type == "cpu"
    assert "mlp_params" in dict(m.named_parameters()) and "bg_color" in dict(m.named_buffers())
    with pytest.raises(NotImplementedError):
        lp.LightplaneRenderer(8, 3, 16, 32, use_naive_impl=True)
    with pytest.raises(ValueError):
        lp.LightplaneRenderer(8, 3, 16, 32, enable_direction_dependent_colors=False)  # harmonics still set
    # encoding consistency rules (renderer_module.py:604-667)
    with pytest.raises(ValueError):
        _check_renderer_ray_encoding_input(torch.zeros(4, 32), 3, 32, True)   # both given
    with pytest.raises(ValueError):
        _check_renderer_ray_encoding_input(None, None, 32, True)               # neither given
    with pytest.raises(ValueError):
        _check_renderer_ray_encoding_input(torch.zeros(4, 8), None, 32, True)  # wrong dim
    _check_renderer_ray_encoding_input(None, 3, 32, True)
    _check_renderer_ray_encoding_input(torch.zeros(4, 32), None, 32, True)
    # CPU tensors never reach a kernel: loud failure, no fallback
    with pytest.raises(RuntimeError):
        m(_rays(4), [torch.zeros(1, 4, 4, 4, 16)])


def test_splatter_modules_and_errors():
    s = lp.LightplaneSplatter(num_samples=8, grid_chn=16)
    assert s.get_splatter_params() is None
    with pytest.raises(ValueError):
        s(_rays(4), [(1, 4, 4, 4, 16)])  # no encoding
    with pytest.raises(ValueError):
        s(_rays(4, torch.zeros(4, 8)), [(1, 4, 4, 4, 16)])  # wrong encoding dim
    ms = lp.LightplaneMLPSplatter(num_samples=8, grid_chn=32, input_grid_chn=16, mlp_hidden_chn=64, mlp_n_layers=3)
    assert ms.get_splatter_params().n_hidden.tolist() == [16, 64, 64, 32]
    with pytest.raises(RuntimeError):
        s(_rays(4, torch.zeros(4, 16)), [(1, 4, 4, 4, 16)])  # CPU tensors: no fallback


def test_cabi_marshalling():
    from lightplane_b200 import _cabi

    gl = _cabi.make_grid_list(torch.zeros(2 * 24 + 2 * 20, 16), [[2, 1, 6, 4, 16], [2, 5, 1, 4, 16]])
    assert gl.num_grids == 2 and gl.channels == 16 and list(gl.sizes[1]) == [2, 5, 1, 4, 16]
    with pytest.raises(AssertionError):
        _cabi.make_grid_list(torch.zeros(10, 16), [[2, 1, 6, 4, 16]])
    cfg = _cabi.make_cfg(128, 4, 2.0, 1e-5, True, False, 0.5, 2**31 + 5, 35)
    assert cfg.noise_num_rays == 48 and cfg.noise_seed == -(2**31) + 5 and cfg.inject_noise == 1
    assert _cabi.make_cfg(8).inject_noise == 0
