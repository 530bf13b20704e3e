"""Stage a runnable copy of the REFERENCE under baseline/_ref/ so that its own Triton kernels can be
run on the GPU box as the GPU comparator.

baseline/_ref/ is git-ignored (the reference's sources never enter this repository's history) but
is shipped to the GPU box by gpurun.  What is staged:
  baseline/_ref/lightplane/   verbatim copy of /root/reference/lightplane with ONE edit:
                              `_floor(x) = x - x % 1` -> `tl.floor(x)` (grid_sample_util.py:12-14);
                              under the installed Triton 3.6 float `%` is C fmod and truncates negative
                              coordinates; the reference pins triton==2.1.0 (floor-mod);
                              and `tl.view(` -> `tl.reshape(` (Triton 3.x `view` may reorder elements when
                              compiled; 2.1.0 did not)
  baseline/_ref/tests/        verbatim copy of the reference's own tests
  baseline/_ref/cogapp, plotly   the stand-ins of oracle/_refshim (neither package is installed)
  baseline/_ref/STAGED.json   provenance

Runs only in the build container (needs /root/reference).  Called by __graft_entry__.build().
"""

import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")


def stage(force=False):
    if not os.path.isdir(REF):
        return False
    marker = os.path.join(DST, "STAGED.json")
    if os.path.exists(marker) and not force:
        return True
    if os.path.exists(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, "lightplane"), os.path.join(DST, "lightplane"))
    shutil.copytree(os.path.join(REF, "tests"), os.path.join(DST, "tests"))
    p = os.path.join(DST, "lightplane", "triton_src", "shared", "grid_sample_util.py")
    src = open(p).read()
    assert "return x - x % 1" in src
    open(p, "w").write(src.replace("return x - x % 1", "return tl.floor(x)"))
    # (2) `tl.view` kept element order under the pinned triton==2.1.0; under Triton 3.x it is
    # reshape(can_reorder=True) and the COMPILED kernels permute rays/channels (the interpreter does
    # not): GPU outputs came out ~50 % off the reference's own naive path.  `tl.reshape` is the
    # order-preserving spelling of what the reference meant.
    n_view = 0
    for root, _, files in os.walk(os.path.join(DST, "lightplane", "triton_src")):
        for f in files:
            if f.endswith(".py"):
                q = os.path.join(root, f)
                t = open(q).read()
                if "tl.view(" in t:
                    n_view += t.count("tl.view(")
                    open(q, "w").write(t.replace("tl.view(", "tl.reshape("))
    for shim in ("cogapp", "plotly"):
        shutil.copytree(os.path.join(REPO, "oracle", "_refshim", shim), os.path.join(DST, shim),
                        ignore=shutil.ignore_patterns("__pycache__"))
    json.dump({"source": REF, "edits": ["grid_sample_util.py: _floor -> tl.floor (SURVEY.md H2)",
                         f"triton_src: tl.view -> tl.reshape ({n_view} call sites; Triton 3.x view may reorder)"],
               "shims": ["cogapp", "plotly"]}, open(marker, "w"))
    return True


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv)
    print("staged" if ok else "no /root/reference here; nothing staged", DST)
