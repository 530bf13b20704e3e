#!/usr/bin/env python
"""Headline benchmark: forward+backward rays/s of the Lightplane Renderer at 128 samples on a
64^2 x 16-channel triplane (BASELINE.json `metric`), synthetic 1920x1080 ray batches.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # CPU arm: the oracle port of the
                                                           # reference's naive path, on host cores

A "step" = one forward + backward pass of the public API (`LightplaneRenderer` module, MSE loss to
a random target image) over the whole ray batch; at N>1 every rank renders its own FullHD batch
(weak scaling) and the grid / MLP gradients are SUM-all-reduced over NCCL inside the step.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how every field is derived.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

METRIC = "renderer_fwd_bwd_rays_per_s"
UNIT = "rays/s"
S = 128
C, H, COLOR = 16, 32, 3
PLANE = 64
# algorithmic per-sample costs: MACs fwd = C*H + H*H + H*H + H + H*H + H*3
MAC_FWD = C * H + H * H + (H * H + H) + (H * H + H * COLOR)
FLOP_FWD_PER_SAMPLE = 2 * MAC_FWD            # 7 424
FLOP_BWD_PER_SAMPLE = 2 * FLOP_FWD_PER_SAMPLE  # dX + dW, recompute NOT counted: 14 848
BYTES_PER_RAY_FWD = 36 + 4 * H + 4 + 4 + 4 * COLOR            # rays+enc in, 3 outputs out
BYTES_PER_RAY_BWD = 36 + 4 * H + 4 + 4 * COLOR + 4 + 4 + 4 * COLOR + 4 * H  # + outputs, grads in, g_enc out


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json (sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback of B200_PROFILING.md")


def camera_rays(width, height, seed, device):
    """Pinhole camera on a sphere of radius 2.6 looking at the origin; the [-1,1]^3 volume spans
    most of the image height; near/far bracket the cube."""
    g = torch.Generator().manual_seed(seed)
    ang = float(torch.rand(1, generator=g)) * 6.28318
    elev = 0.3 + 0.4 * float(torch.rand(1, generator=g))
    eye = 2.6 * torch.tensor([torch.cos(torch.tensor(ang)) * torch.cos(torch.tensor(elev)),
                              torch.sin(torch.tensor(elev)),
                              torch.sin(torch.tensor(ang)) * torch.cos(torch.tensor(elev))])
    fwd = -eye / eye.norm()
    up = torch.tensor([0.0, 1.0, 0.0])
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    aspect = width / height
    ys, xs = torch.meshgrid(torch.linspace(0.62, -0.62, height), torch.linspace(-0.62 * aspect, 0.62 * aspect, width),
                            indexing="ij")
    dirs = fwd[None, None] + xs[..., None] * right + ys[..., None] * up
    dirs = (dirs / dirs.norm(dim=-1, keepdim=True)).reshape(-1, 3).contiguous()
    n = dirs.shape[0]
    origins = eye.expand(n, 3).contiguous()
    near = torch.full((n,), 2.6 - 1.75)
    far = torch.full((n,), 2.6 + 1.75)
    grid_idx = torch.zeros(n, dtype=torch.int32)
    return dirs, origins, grid_idx, near, far


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference's naive PyTorch path) on the host cores
# ---------------------------------------------------------------------------------------------
def oracle_fwd_bwd(n_rays, chunk, seed=0):
    """fwd+bwd of `n_rays` rays of the bench workload through oracle/lightplane_oracle.py, chunked
    over rays (the naive formulation keeps every per-sample activation); returns seconds."""
    from oracle import lightplane_oracle as O

    import lightplane_b200 as lp

    torch.manual_seed(seed)
    dp = lp.init_decoder_params("cpu", 2, 2, 2, input_chn=C, hidden_chn=H, color_chn=COLOR, opacity_init_bias=-1.0)
    shapes = [[1, 1, PLANE, PLANE, C], [1, PLANE, 1, PLANE, C], [1, PLANE, PLANE, 1, C]]
    rows = sum(s[1] * s[2] * s[3] for s in shapes)
    grid = (0.5 * torch.randn(rows, C)).requires_grad_(True)
    mlp = dp.mlp_params.clone().requires_grad_(True)
    side = int(n_rays ** 0.5)
    d, o, gi, nr, fr = camera_rays(side, n_rays // side, seed, "cpu")
    n = d.shape[0]
    enc = torch.randn(n, H)
    target = torch.rand(n, COLOR)
    t0 = time.perf_counter()
    for lo in range(0, n, chunk):
        sl = slice(lo, min(lo + chunk, n))
        _, _, feat = O.render(d[sl], o[sl], gi[sl].long(), nr[sl], fr[sl], enc[sl], grid, shapes, mlp,
                              [C, H, H], [H, H, 1], [H, H, 16], num_samples=S, gain=1.0)
        loss = ((feat[:, :COLOR] - target[sl]) ** 2).sum()
        loss.backward()
    return time.perf_counter() - t0, n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays, chunk = 8192, 2048
    threads = torch.get_num_threads()
    for _ in range(args.warmup):
        oracle_fwd_bwd(2048, 2048)
    t = 0.0
    n = 0
    for k in range(args.steps):
        dt, nn = oracle_fwd_bwd(n_rays, chunk, seed=k)
        t += dt
        n += nn
    value = n / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"renderer fwd+bwd, {S} samples, triplane {PLANE}^2x{C}ch, MLP 2/2/2 h{H}",
                   "note": "reference is a Python package (no compiled CPU path): this arm is the oracle "
                           "port of its naive PyTorch renderer, chunked over rays, on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_rays} rays x {S} samples per step, chunks of {chunk}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-tile-walk", dest="tile_walk", action="store_false",
                    help="do not pass the image width (ray_image_width): rays are walked as 128-ray scan-line runs")
    # non-default workloads (e.g. BASELINE.json configs[4]: --samples 256 --plane 128 --grid-chn 32)
    ap.add_argument("--samples", type=int, default=S)
    ap.add_argument("--plane", type=int, default=PLANE)
    ap.add_argument("--grid-chn", type=int, default=C)
    args = ap.parse_args()
    globals().update(S=args.samples, PLANE=args.plane, C=args.grid_chn)
    globals().update(MAC_FWD=C * H + H * H + (H * H + H) + (H * H + H * COLOR))
    globals().update(FLOP_FWD_PER_SAMPLE=2 * MAC_FWD, FLOP_BWD_PER_SAMPLE=4 * MAC_FWD)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist

    import lightplane_b200 as lp
    from lightplane_b200 import _cabi
    from lightplane_b200.distributed import all_reduce_gradients

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- model (replicated) + data (per rank) ----
    torch.manual_seed(0)
    model = lp.LightplaneRenderer(num_samples=S, color_chn=COLOR, grid_chn=C, mlp_hidden_chn=H,
                                  opacity_init_bias=-1.0).to(dev)
    shapes = [[1, 1, PLANE, PLANE, C], [1, PLANE, 1, PLANE, C], [1, PLANE, PLANE, 1, C]]
    grids = [(0.5 * torch.randn(s, device=dev)).requires_grad_(True) for s in shapes]
    params = grids + list(model.parameters())
    host = [t.pin_memory() for t in camera_rays(args.width, args.height, 1000 + rank, "cpu")]
    n_rays = host[0].shape[0]
    target_host = torch.rand(n_rays, COLOR, generator=torch.Generator().manual_seed(rank)).pin_memory()
    resident = [t.to(dev) for t in host]
    target = target_host.to(dev)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host) + target_host.numel() * 4
    out_host = torch.empty(n_rays, COLOR).pin_memory()
    loss_host = torch.empty(1).pin_memory()
    d2h_bytes = out_host.numel() * 4 + 4

    def step(rays_t, tgt):
        for p in params:
            p.grad = None
        rays = lp.Rays(directions=rays_t[0], origins=rays_t[1], grid_idx=rays_t[2], near=rays_t[3], far=rays_t[4])
        _, _, feat = model(rays, grids, ray_image_width=args.width if args.tile_walk else None)
        loss = ((feat - tgt) ** 2).mean()
        loss.backward()
        if world > 1:
            all_reduce_gradients(params)
        return feat, loss

    def e2e_step():
        rays_t = [t.to(dev, non_blocking=True) for t in host]
        tgt = target_host.to(dev, non_blocking=True)
        feat, loss = step(rays_t, tgt)
        out_host.copy_(feat.detach(), non_blocking=True)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        step(resident, target)
        e2e_step()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    _cabi.profile_begin()
    ms = timed(lambda: step(resident, target), args.steps)
    launches = _cabi.profile_end()
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(e2e_step, args.steps)

    value = world * n_rays * args.steps / (ms * 1e-3)
    e2e_value = world * n_rays * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = load_peaks()
        per = {}
        for name, t in launches:
            per.setdefault(name, []).append(t)
        avg = {k: sum(v) / len(v) for k, v in per.items()}
        bwd_ms = avg.get("lp_render_backward", float("nan"))
        fwd_ms = avg.get("lp_render_forward", float("nan"))
        samples = n_rays * S
        ach_tflops = samples * FLOP_BWD_PER_SAMPLE / (bwd_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("lp_render_backward_dram_bytes_per_launch")
        roofline = {
            "kernel": "lp_render_backward", "bound": "tensor", "achieved": ach_tflops, "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": ach_tflops / peaks["tflops"], "traffic": traffic,
            "peak_source": peaks["source"],
            "algorithmic_flops_per_launch": samples * FLOP_BWD_PER_SAMPLE,
            "launch_ms": {"lp_render_forward": fwd_ms, "lp_render_backward": bwd_ms},
            "kernel_share_of_step": (fwd_ms + bwd_ms) * args.steps / ms,
            "hbm": {"algorithmic_bytes_per_launch": n_rays * BYTES_PER_RAY_BWD,
                    "achieved_gbs": n_rays * BYTES_PER_RAY_BWD / (bwd_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks["hbm_gbs"],
                    "frac": n_rays * BYTES_PER_RAY_BWD / (bwd_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "forward": {"achieved_tflops": samples * FLOP_FWD_PER_SAMPLE / (fwd_ms * 1e-3) / 1e12},
        }
        cpu = None
        if world == 1 and not args.skip_cpu_baseline:
            dt, nn = oracle_fwd_bwd(8192, 2048)
            dt, nn = oracle_fwd_bwd(16384, 2048, seed=1)
            cpu = {"value": nn / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": "16384 rays x 128 samples of the same workload, chunks of 2048 rays, 1 warm-up"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{args.width}x{args.height} rays/GPU, {S} samples, triplane 3x[{PLANE}x{PLANE}]x{C}ch, "
                            f"MLP trunk/opacity/colour 2/2/2 hidden {H}, colour {COLOR}; fwd+bwd of LightplaneRenderer + MSE",
                "rays_per_gpu": n_rays, "num_samples": S,
                "l2_policy": "inputs larger than L2: per-step ray/encoding/gradient tensors ~%d MB" % (n_rays * (BYTES_PER_RAY_BWD) // 1e6),
                "parallelism": f"rays sharded x{world}, grid+MLP replicated, grad all-reduce (NCCL)" if world > 1 else "single GPU",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": len(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
