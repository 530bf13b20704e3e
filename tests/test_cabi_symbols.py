"""CPU checks of the boundary: the built library exports every symbol the header declares and the
binding's prototypes cover the header; no compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from lightplane_b200 import _cabi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "lightplane_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lp_[a-z_0-9]+)\s*\(", src)))


def test_binding_covers_header():
    assert sorted(_cabi.EXPORTED_SYMBOLS) == _declared_symbols()


def test_device_library_exports_all_symbols():
    if not os.path.exists(_cabi.DEFAULT_LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(REPO, "lightplane_b200", "csrc")], check=True)
    lib = ctypes.CDLL(_cabi.DEFAULT_LIB_PATH)  # links against libcudart only; loads without a GPU
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.lp_abi_version.restype = ctypes.c_int
    assert lib.lp_abi_version() == _cabi.LP_ABI_VERSION
    assert lib.lp_is_device_build() == 1


def test_struct_sizes_match_header():
    """ctypes mirrors of the C structs must have the C layout (checked against a tiny C program)."""
    prog = r'''
#include <stdio.h>
#include "lightplane_b200.h"
int main(){printf("%zu %zu %zu %zu %zu\n", sizeof(lp_grid_list), sizeof(lp_rays), sizeof(lp_march_cfg),
 sizeof(lp_decoder_spec), sizeof(lp_mlp_spec)); return 0;}
'''
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), "-o", os.path.join(td, "s"),
                        os.path.join(td, "s.c")], check=True)
        out = subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout
    sizes = [int(v) for v in out.split()]
    assert sizes == [ctypes.sizeof(_cabi.GridList), ctypes.sizeof(_cabi.RaysStruct),
                     ctypes.sizeof(_cabi.MarchCfg), ctypes.sizeof(_cabi.DecoderSpec),
                     ctypes.sizeof(_cabi.MlpSpec)]


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.LightplaneB200Error):
        _cabi.load_library(str(tmp_path / "nope.so"))
