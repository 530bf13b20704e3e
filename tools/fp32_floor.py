#!/usr/bin/env python
"""Error of the tensor-core kernels AND of the generic fp32 kernels (LP_ONLY_GENERIC=1) against the fp64 oracle on the
same problems: what a plain fp32 implementation of the decoder achieves is the floor the fast paths' gradient
tolerances are judged against.  Prints a markdown table (profiles/fp32_floor_r2.md)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import torch  # noqa: E402

from _golden import coherent_case, oracle_render_case, rel_err, synthetic_case  # noqa: E402
from _lowlevel import render_case  # noqa: E402
from lightplane_b200 import _cabi  # noqa: E402

CASES = [("2/2/2 C16", dict(layers=(2, 2, 2), C=16, n=1500), None),
         ("4/2/4 C16", dict(layers=(4, 2, 4), C=16, n=1500), None),
         ("2/4/2 C32 scaffold", dict(layers=(2, 4, 2), C=32, n=700), 10),
         ("1/1/1 C16", dict(layers=(1, 1, 1), C=16, n=900), None),
         ("3/1/2 C32", dict(layers=(3, 1, 2), C=32, n=600), None),
         ("1/3/1 C16", dict(layers=(1, 3, 1), C=16, n=500), None)]


def main():
    lib = _cabi.get_lib()
    keys = ("features", "g_grid", "g_mlp", "g_enc")
    print("| decoder | path | " + " | ".join(keys) + " |")
    print("|---|---|" + "---|" * len(keys))
    for name, kw, scaf in CASES:
        c = synthetic_case(hidden=32, color_grid=False, plane=40, samples=24, samples_inf=3, pixel=0.004, batch=1, **kw)
        if scaf:
            c = coherent_case(c, n=kw["n"], pixel=0.004, seed=3, scaffold_res=scaf)
        want = oracle_render_case(c)
        for path, env in (("tensor core", "0"), ("generic fp32", "1")):
            os.environ["LP_ONLY_GENERIC"] = env
            got = render_case(lib, c, "cuda")
            print(f"| {name} | {path} | " + " | ".join(f"{rel_err(got[k], want[k]):.1e}" for k in keys) + " |")
    # the bench camera (configs[1]/[2] shape: 128 samples, 64^2 x 16 triplane), both losses of SURVEY.md 8d
    import test_gpu_baseline_configs as T
    for loss, tile in (("randsign", False), ("mse", False)):
        for path, env in (("tensor core", "0"), ("generic fp32", "1")):
            os.environ["LP_ONLY_GENERIC"] = env
            errs = T.bench_camera_errors(loss, tile)
            print(f"| bench camera 4096x128, {loss} | {path} | " + " | ".join(f"{errs[k]:.1e}" for k in keys) + " |")
    os.environ["LP_ONLY_GENERIC"] = "0"


if __name__ == "__main__":
    main()
