"""GPU-less debugging aid: the CUDA kernel sources compiled for the host against the SIMT
emulation shim (tests/hostsim/) are run on the golden cases and compared with the reference's
outputs.  This exercises kernel LOGIC only; the parity tests proper are the `-m gpu` tests."""
import os
import subprocess

import pytest
import torch

from _golden import (case_names, coherent_case, load_case, oracle_render_case, oracle_splat_case, rel_err, plain_splat_case, regrid_case,
                     synthetic_case, synthetic_splat_case)
from _lowlevel import render_case, splat_case
from lightplane_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "hostsim", "liblp_hostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "lightplane_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", CSRC, "hostsim"], check=True)
    lib = _cabi.load_library(LIB)
    assert lib.lp_is_device_build() == 0
    return lib


@pytest.mark.parametrize("name", case_names("render_"))
def test_hostsim_renderer(lib, name):
    c = load_case(name)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        assert torch.isfinite(v).all(), (name, k)
        # tensor-core path (2/2/2 x hidden 32 cases): TF32 backward products, bf16 dW operands
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, c["naive_" + k]) < tol, (name, k, rel_err(v, c["naive_" + k]))


@pytest.mark.parametrize("name", case_names("splat_"))
def test_hostsim_splatter(lib, name):
    c = load_case(name)
    got = splat_case(lib, c, "cpu")
    for k, v in got.items():
        assert torch.isfinite(v).all(), (name, k)
        assert rel_err(v, c["naive_" + k]) < 2e-4, (name, k, rel_err(v, c["naive_" + k]))


@pytest.mark.parametrize("name,pixel,mask,scaf", [("render_triplane_inf_gain", 0.02, 0, None), ("render_triplane_inf_gain", 0.08, 1, None),
                                                  ("render_c32_b1", 0.03, 1, None), ("render_triplane_inf_gain", 0.02, 0, 8),
                                                  ("render_c32_b1", 0.03, 1, 6)])
def test_hostsim_renderer_coherent_rays(lib, name, pixel, mask, scaf):
    """Neighbouring-pixel rays (overlapping footprints, many samples outside the planes), unlike the
    golden cases' random rays."""
    c = coherent_case(load_case(name), n=64, pixel=pixel, mask_oob=mask, scaffold_res=scaf)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (name, k, rel_err(v, want[k]))


@pytest.mark.parametrize("C,sigma", [(16, 0.0), (32, 0.5)])
def test_hostsim_renderer_color_grid_tensor_core_path(lib, C, sigma):
    """Separate colour grid ("ReLU field", trunk-less decoder, hidden 32): lp_render_tc_cg.cuh."""
    c = synthetic_case(n=96, C=C, sigma=sigma)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (C, k, rel_err(v, want[k]))


@pytest.mark.parametrize("c_in,c_out,triplane", [(16, 16, False), (32, 16, True), (16, 32, True)])
def test_hostsim_mlp_splatter_tensor_core_path(lib, c_in, c_out, triplane):
    """MLP splatter [c_in -> 32 -> c_out] on its tensor-core path (lp_splat_tc.cuh)."""
    c = synthetic_splat_case(n=96, c_in=c_in, c_out=c_out, triplane=triplane)
    want = oracle_splat_case(c)
    got = splat_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else 2e-4
        assert torch.isfinite(v).all(), (k,)
        assert rel_err(v, want[k]) < tol, (c_in, c_out, k, rel_err(v, want[k]))


@pytest.mark.parametrize("name,mask", [("render_triplane_inf_gain", 0), ("render_c32_b1", 1)])
def test_hostsim_renderer_empty_space_folding(lib, name, mask):
    """Rays that start and end far outside the volume: at many steps all samples of a group miss every plane, the
    case the tensor-core kernels fold (decoder evaluated once at zero features, summed compositing gradients)."""
    c = coherent_case(load_case(name), n=160, pixel=0.01, mask_oob=mask, origin=(1.3, -0.2, -3.0), near=0.3, far=6.0)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (name, k, rel_err(v, want[k]))


@pytest.mark.parametrize("C,sigma,scaf", [(32, 0.0, None), (16, 0.5, 6)])
def test_hostsim_renderer_hidden64_forward(lib, C, sigma, scaf):
    """Hidden width 64 (the reference's example configuration): lp_render_tc_wide.cuh."""
    c = synthetic_case(n=96, C=C, hidden=64, layers=(2, 2, 2), color_grid=False, sigma=sigma)
    if scaf:
        c = coherent_case(c, n=96, pixel=0.03, seed=3, scaffold_res=scaf)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (C, k, rel_err(v, want[k]))


@pytest.mark.parametrize("layers,C,sigma", [((4, 2, 4), 16, 0.0), ((2, 4, 2), 32, 0.5), ((1, 1, 1), 16, 0.0), ((3, 1, 2), 32, 0.0),
                                            ((1, 3, 1), 16, 0.0)])
def test_hostsim_renderer_layer_counts_tensor_core_path(lib, layers, C, sigma):
    """Layer counts other than 2/2/2 (hidden 32): the table-driven tensor-core kernels of lp_render_tc_deep.cuh."""
    c = synthetic_case(n=96, C=C, hidden=32, layers=layers, color_grid=False, sigma=sigma)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (layers, k, rel_err(v, want[k]))


def test_hostsim_layer_counts_empty_space_folding(lib):
    c = synthetic_case(n=160, C=16, hidden=32, layers=(3, 2, 3), color_grid=False)
    c = coherent_case(c, n=160, pixel=0.01, seed=3, mask_oob=1, origin=(1.3, -0.2, -3.0), near=0.3, far=6.0)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (k, rel_err(v, want[k]))


def test_hostsim_renderer_tile_walk_hint(lib):
    """`ray_image_width`: the same rays walked as 16x8-pixel tiles give the same results up to summation order (the
    MMAs of a product accumulate in issue order, gradient reductions in arrival order)."""
    c = coherent_case(load_case("render_triplane_inf_gain"), n=256, pixel=0.02, mask_oob=0)
    a = render_case(lib, c, "cpu")
    b = render_case(lib, c, "cpu", ray_image_width=16)
    for k in a:
        assert rel_err(a[k], b[k]) < (1e-4 if k != "g_mlp" else 2e-3), (k, rel_err(a[k], b[k]))


@pytest.mark.parametrize("hidden", [32, 16])
def test_hostsim_single_sample_with_background_samples(lib, hidden):
    """num_samples == 1 with num_samples_inf > 0: the first background sample's step length is measured from `far`
    (depth_inv_sphere(..., -1) == far), on the tensor-core path (hidden 32) and on the generic path (hidden 16)."""
    c = synthetic_case(n=64, C=16, hidden=hidden, layers=(2, 2, 2), color_grid=False, samples=1, samples_inf=3)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (hidden, k, rel_err(v, want[k]))


def test_hostsim_renderer_several_tiles_per_group(lib):
    """More ray tiles than groups (the emulated device has 2 SMs x 2 groups): every group walks several tiles in
    sequence -- the tile tail, the next tile's per-ray constant and probe slot, and the memory group running ahead
    across the tile boundary in the warp-specialised backward."""
    c = coherent_case(load_case("render_triplane_inf_gain"), n=1100, pixel=0.004, mask_oob=0)
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu", ray_image_width=0)
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (k, rel_err(v, want[k]))


@pytest.mark.parametrize("channels,samples,triplane,mask", [(4, 5, False, 1), (8, 7, True, 0), (32, 13, False, 1), (64, 9, True, 1),
                                                            (128, 6, False, 0), (256, 40, True, 1)])
def test_hostsim_plain_splatter_shared_march(lib, channels, samples, triplane, mask):
    """Sub-warps of 1..32 lanes per ray (lp_splat.cuh: one lane per sample works out the taps, shuffles hand them round),
    sample counts that are not multiples of the sub-warp width, a ray count that leaves sub-warps of the last warp idle."""
    c = plain_splat_case(n=70, channels=channels, samples=samples, samples_inf=3, triplane=triplane, mask_oob=mask, batch=2)
    want = oracle_splat_case(c)
    got = splat_case(lib, c, "cpu")
    for k, v in got.items():
        assert torch.isfinite(v).all(), k
        assert rel_err(v, want[k]) < 2e-4, (k, rel_err(v, want[k]))


@pytest.mark.parametrize("variant", ["order_yz_xy_xz", "order_xz_yz_xy", "sizes_not_one_volume"])
def test_hostsim_triplane_fast_path_order_and_fallback(lib, variant):
    """The triplane fast path (LpGridSet::tri) for plane lists in any order, and the generic path for three planes whose
    axis sizes do not agree."""
    c = synthetic_case(n=128, C=16, hidden=32, layers=(2, 2, 2), color_grid=False, plane=7, samples=9, samples_inf=2, pixel=0.05)
    if variant == "order_yz_xy_xz":
        c = regrid_case(c, order=(2, 0, 1))
    elif variant == "order_xz_yz_xy":
        c = regrid_case(c, order=(1, 2, 0))
    else:
        c = regrid_case(c, sizes=[[2, 1, 7, 8, 16], [2, 5, 1, 9, 16], [2, 6, 7, 1, 16]])
    want = oracle_render_case(c)
    got = render_case(lib, c, "cpu")
    for k, v in got.items():
        tol = 6e-3 if k == "g_mlp" else (1e-3 if k.startswith("g_") else 2e-4)
        assert rel_err(v, want[k]) < tol, (variant, k, rel_err(v, want[k]))
