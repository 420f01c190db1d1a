"""CPU-side checks of the product library: it loads, exports every symbol the header declares,
and refuses to create a handle without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT
from sph_project_b200 import _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)))


def test_binding_covers_header():
    assert header_symbols() == sorted(_native.PROTOTYPES)


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_library_exports_every_symbol(which):
    if which == "cuda":
        if not os.path.exists(_native.CUDA_LIBRARY_PATH):
            import __graft_entry__ as g
            g.build()
        lib = ctypes.CDLL(_native.CUDA_LIBRARY_PATH)
    else:
        from helpers import oracle_library
        lib = oracle_library()
    for name in header_symbols():
        assert hasattr(lib, name), f"{which} library lacks {name}"
    _native.bind(lib)
    assert lib.sph_abi_version() == _native.ABI_VERSION
    assert lib.sph_backend_name().decode() == {"cuda": "cuda-sm100a", "oracle": "oracle-cpu"}[which]


def test_enum_values_match_header():
    text = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    enums = dict(re.findall(r"\b(SPH_[FST]_[A-Z0-9_]+)\s*=\s*(\d+)", text))
    for prefix, cls in (("SPH_F_", _native.F), ("SPH_S_", _native.S), ("SPH_T_", _native.T)):
        names = {k[len(prefix):]: int(v) for k, v in enums.items() if k.startswith(prefix)}
        mine = {k: v for k, v in vars(cls).items() if k.isupper()}
        assert names == mine, prefix


def test_struct_sizes():
    # the C compiler's layout of the ABI structs (checked against a tiny C program)
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "sph_b200.h"\nint main(){printf("%zu %zu %zu", sizeof(SphParams), sizeof(SphStepStats), sizeof(SphSlabInfo));}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [ctypes.sizeof(_native.SphParams), ctypes.sizeof(_native.SphStepStats), ctypes.sizeof(_native.SphSlabInfo)]


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import make_sim, scene
    with pytest.raises(_native.SphError):
        make_sim(scene("wcsph"))
