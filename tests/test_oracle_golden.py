"""Pins the oracle (oracle/lightplane_oracle.py) against the golden vectors produced by the
reference's own code (oracle/make_golden.py): its naive PyTorch path and its Triton kernels run
under TRITON_INTERPRET=1.  CPU only."""
import pytest
import torch

from _golden import (case_names, load_case, oracle_render_case, oracle_splat_case, rel_err)

RENDER_KEYS = ("ray_length", "nlt", "features", "g_grid", "g_mlp", "g_enc")


@pytest.mark.parametrize("name", case_names("render_"))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_renderer_oracle_matches_reference(name, dtype):
    c = load_case(name)
    got = oracle_render_case(c, dtype=dtype)
    tol = 2e-5 if dtype == torch.float64 else 5e-5
    keys = RENDER_KEYS + (("g_color_grid",) if "color_grid" in c else ())
    has_inf = int(c["cfg"][1]) > 0
    noisy_pad = float(c["cfg_f"][2]) > 0 and c["directions"].shape[0] % 16 != 0
    for k in keys:
        # naive path of the reference: fp32 on CPU
        assert rel_err(got[k], c["naive_" + k]) < tol, (name, k, "naive")
        # Triton kernels of the reference (interpreted).  With background samples they evaluate
        # the disparity schedule in fp32 and are themselves ~1e-3..1e-2 off the naive path
        # (see oracle/make_golden.py), so only a loose bound applies there.
        if noisy_pad:
            continue
        t_tol = 2e-2 if has_inf else tol
        assert rel_err(got[k], c["triton_" + k]) < t_tol, (name, k, "triton")


@pytest.mark.parametrize("name", case_names("splat_"))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_splatter_oracle_matches_reference(name, dtype):
    c = load_case(name)
    got = oracle_splat_case(c, dtype=dtype)
    tol = 2e-5 if dtype == torch.float64 else 5e-5
    for k in got:
        assert rel_err(got[k], c["naive_" + k]) < tol, (name, k, "naive")
        assert rel_err(got[k], c["triton_" + k]) < tol, (name, k, "triton")


def test_rng_matches_reference_distribution():
    from oracle.lightplane_oracle import int_to_randn

    x1 = torch.arange(1, 200001)
    z = int_to_randn(x1, x1 + 12345, seed=3)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
