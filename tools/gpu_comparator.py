"""GPU comparator: run the REFERENCE's own Triton kernels on the same
B200, on the same tensors, next to this repository's CUDA path.

  * parity GPU-vs-GPU: outputs and gradients of `lightplane_renderer` (reference, Triton, staged copy
    under baseline/_ref with the one-line `_floor` fix of SURVEY.md H2) vs `lightplane_b200`,
    mean|d|/mean|ref| per tensor, for camera rays, border-crossing rays (the reference's own test
    generator) and the `num_samples_inf` case the golden fixtures exclude;
  * timing: forward+backward rays/s of both on BASELINE.json configs[1] (256x256) and configs[2]
    (1920x1080), CUDA events, >= 5 timed after >= 2 warm-ups (per-constexpr JIT warm-up included).

Writes gpurun_out/gpu_comparator.json.  Test/measurement infrastructure: nothing in the product path
imports baseline/_ref.  Usage on the GPU box:  python tools/gpu_comparator.py [--quick]
"""
import argparse
import json
import os
import sys
import time
import traceback

# The reference pins triton==2.1.0; under the installed Triton 3.6 its kernels read module-level Python globals
# (ALLOW_TF32 etc., triton_src/shared/const.py), which newer Triton only permits with this switch.  No source edit.
os.environ.setdefault("TRITON_ALLOW_NON_CONSTEXPR_GLOBALS", "1")

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(REPO, "baseline", "_ref")
sys.path.insert(0, REPO)

import torch  # noqa: E402

OUT = os.path.join(REPO, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def errstr(ex):
    parts = []
    while ex is not None and len(parts) < 4:
        parts.append("".join(traceback.format_exception_only(type(ex), ex)).strip()[-700:])
        ex = ex.__cause__ or ex.__context__
    return " <- ".join(parts)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().mean() / b.abs().mean().clamp_min(1e-30))


def relmax(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def camera(width, height, seed, dev):
    from bench import camera_rays
    return [t.to(dev) for t in camera_rays(width, height, seed, "cpu")]


def border_rays(n, seed, dev):
    """the reference's test generator (tests/utils.py:230-268): rays cross the cube's borders"""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g) / 3.0
    d = -o + 0.1 * torch.randn(n, 3, generator=g)
    near = torch.randn(n, generator=g) * 0.1 + 0.1
    far = torch.randn(n, generator=g).abs() * 0.1 + 3.0
    gi = torch.zeros(n, dtype=torch.int32)
    return [t.to(dev) for t in (d, o, gi, near, far)]


def make_problem(dev, C=16, H=32, plane=64, seed=0):
    import lightplane_b200 as lp
    torch.manual_seed(seed)
    dp = lp.init_decoder_params(dev, 2, 2, 2, input_chn=C, hidden_chn=H, color_chn=3, opacity_init_bias=-1.0)
    shapes = [[1, 1, plane, plane, C], [1, plane, 1, plane, C], [1, plane, plane, 1, C]]
    grids = [0.5 * torch.randn(s, device=dev) for s in shapes]
    return dp, grids


def run_impl(mod, DP, dp, grids, rays_t, enc, cot, loss_kind, **kw):
    """one forward+backward through module `mod` (reference or ours); returns outputs + grads"""
    gl = [g.detach().clone().requires_grad_(True) for g in grids]
    mp = dp.mlp_params.detach().clone().requires_grad_(True)
    e = enc.detach().clone().requires_grad_(True)
    d = DP(mp, dp.n_hidden_trunk, dp.n_hidden_opacity, dp.n_hidden_color, dp.color_chn)
    rays = mod.Rays(directions=rays_t[0], origins=rays_t[1], grid_idx=rays_t[2], near=rays_t[3], far=rays_t[4],
                    encoding=e)
    kw = dict(gain=1.0) | kw
    outs = mod.lightplane_renderer(rays, gl, d, **kw)
    if loss_kind == "randsign":   # the reference tests' loss (tests/test_renderer_with_autograd.py:211-213)
        loss = sum((c * v).sum() for c, v in zip(cot, outs))
    else:                          # MSE image loss to a random target
        loss = ((outs[2] - cot[2]) ** 2).mean()
    loss.backward()
    C = grids[0].shape[-1]
    return dict(ray_length=outs[0].detach(), nlt=outs[1].detach(), features=outs[2].detach(),
                g_grid=torch.cat([g.grad.reshape(-1, C) for g in gl], 0), g_mlp=mp.grad, g_enc=e.grad)


def time_impl(mod, DP, dp, grids, rays_t, enc, tgt, warm, reps, **kw):
    gl = [g.detach().clone().requires_grad_(True) for g in grids]
    mp = dp.mlp_params.detach().clone().requires_grad_(True)
    e = enc.detach().clone().requires_grad_(True)
    d = DP(mp, dp.n_hidden_trunk, dp.n_hidden_opacity, dp.n_hidden_color, dp.color_chn)
    rays = mod.Rays(directions=rays_t[0], origins=rays_t[1], grid_idx=rays_t[2], near=rays_t[3], far=rays_t[4],
                    encoding=e)

    def step():
        for t in gl + [mp, e]:
            t.grad = None
        outs = mod.lightplane_renderer(rays, gl, d, **kw)
        ((outs[2] - tgt) ** 2).mean().backward()

    t0 = time.perf_counter()
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    jit_s = time.perf_counter() - t0
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return {"ms_median": times[len(times) // 2], "ms_min": times[0], "ms_all": times, "warmup_s": jit_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="skip the FullHD timing of the reference")
    ap.add_argument("--skip-ours", action="store_true")
    args = ap.parse_args()
    res = {"box": {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}, "parity": [], "timing": []}
    if not os.path.isdir(os.path.join(REFDIR, "lightplane")):
        res["unavailable"] = "baseline/_ref not staged (run tools/stage_reference.py in the build container)"
        json.dump(res, open(os.path.join(OUT, "gpu_comparator.json"), "w"), indent=1)
        print(json.dumps(res))
        return
    sys.path.insert(0, REFDIR)
    import triton
    res["box"]["triton"] = triton.__version__
    import lightplane as ref  # the staged reference copy
    assert ref.__file__.startswith(REFDIR), ref.__file__
    import lightplane_b200 as lp
    dev = torch.device("cuda", 0)
    dp, grids = make_problem(dev)

    def save():
        json.dump(res, open(os.path.join(OUT, "gpu_comparator.json"), "w"), indent=1)

    # ---------------- parity on identical tensors ----------------
    cases = [
        ("camera_4096x128_mse", camera(64, 64, 1000, dev), dict(num_samples=128), "mse"),
        ("camera_4096x128_randsign", camera(64, 64, 1000, dev), dict(num_samples=128), "randsign"),
        ("camera_65536x128_mse (configs[1])", camera(256, 256, 1000, dev), dict(num_samples=128), "mse"),
        ("border_4096x64_randsign", border_rays(4096, 3, dev), dict(num_samples=64), "randsign"),
        ("border_4096x48_inf16_randsign", border_rays(4096, 4, dev), dict(num_samples=48, num_samples_inf=16), "randsign"),
        ("border_4096x64_gain3_maskoob", border_rays(4096, 5, dev), dict(num_samples=64, gain=3.0,
                                                                           mask_out_of_bounds_samples=True), "randsign"),
    ]
    for name, rays_t, kw, loss_kind in cases:
        n = rays_t[0].shape[0]
        g = torch.Generator().manual_seed(7)
        enc = torch.randn(n, 32, generator=g).to(dev)
        if loss_kind == "mse":
            cot = [None, None, torch.rand(n, 3, generator=g).to(dev)]
        else:
            cot = [torch.randn(n, generator=g).to(dev), torch.randn(n, generator=g).to(dev),
                   torch.randn(n, 3, generator=g).to(dev)]
        row = {"case": name, "n_rays": n, "kwargs": kw, "loss": loss_kind}
        try:
            r = run_impl(ref, ref.DecoderParams, dp, grids, rays_t, enc, cot, loss_kind, **kw)
            row["reference_ok"] = True
            if n <= 4096:  # the reference's OWN two fp32 paths against each other: the agreement any fp32 implementation can have
                import types
                naive = types.SimpleNamespace(Rays=ref.Rays, lightplane_renderer=ref.lightplane_renderer_naive)
                rn = run_impl(naive, ref.DecoderParams, dp, grids, rays_t, enc, cot, loss_kind, **kw)
                row["reference_naive_vs_triton"] = {k: rel(rn[k], r[k]) for k in r}
                if not args.skip_ours:
                    o_ = run_impl(lp, lp.DecoderParams, dp, grids, rays_t, enc, cot, loss_kind, **kw)
                    row["ours_vs_reference_naive"] = {k: rel(o_[k], rn[k]) for k in r}
            if not args.skip_ours:
                o = run_impl(lp, lp.DecoderParams, dp, grids, rays_t, enc, cot, loss_kind, **kw)
                row["mean_rel"] = {k: rel(o[k], r[k]) for k in r}
                row["max_rel"] = {k: relmax(o[k], r[k]) for k in r}
        except Exception as ex:  # keep going: the report says what failed
            row["error"] = errstr(ex)
        print(json.dumps(row), flush=True)
        res["parity"].append(row)
        save()

    # ---------------- timing: configs[1] and configs[2] ----------------
    sizes = [("configs[1] 256x256", 256, 256)] + ([] if args.quick else [("configs[2] 1920x1080", 1920, 1080)])
    for name, w, h in sizes:
        rays_t = camera(w, h, 1000, dev)
        n = rays_t[0].shape[0]
        g = torch.Generator().manual_seed(9)
        enc = torch.randn(n, 32, generator=g).to(dev)
        tgt = torch.rand(n, 3, generator=g).to(dev)
        row = {"workload": name + " rays, 128 samples, triplane 64^2x16, MLP 2/2/2 h32, fwd+bwd + MSE", "n_rays": n}
        for label, mod, DP in (("reference_triton", ref, ref.DecoderParams), ("lightplane_b200", lp, lp.DecoderParams)):
            if label == "lightplane_b200" and args.skip_ours:
                continue
            try:
                t = time_impl(mod, DP, dp, grids, rays_t, enc, tgt, warm=2, reps=5, num_samples=128, gain=1.0)
                t["rays_per_s"] = n / (t["ms_median"] * 1e-3)
                row[label] = t
            except Exception as ex:
                row[label] = {"error": errstr(ex)}
            print(json.dumps({name: {label: row[label]}}), flush=True)
            save()
        if "rays_per_s" in row.get("reference_triton", {}) and "rays_per_s" in row.get("lightplane_b200", {}):
            row["speedup"] = row["lightplane_b200"]["rays_per_s"] / row["reference_triton"]["rays_per_s"]
        res["timing"].append(row)
        save()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
