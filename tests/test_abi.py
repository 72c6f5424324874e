"""CPU: the C-ABI library loads and exports every symbol include/ssf_b200.h declares, and the ctypes table agrees
with the header (argument count and pointer/int/float kinds).  No compute is called."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    import glob
    text = "".join(open(h).read() for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(ssf_\w+)\s*\(([^)]*)\)\s*;", text):
        args = [a.strip() for a in m.group(3).split(",") if a.strip() and a.strip() != "void"]
        decls[m.group(2)] = args
    return decls


def _kind(arg):
    if "*" in arg:
        return "p"
    t = arg.rsplit(" ", 1)[0].strip()
    return {"int": "i", "float": "f", "double": "d", "unsigned long long": "Q", "long long": "q"}[t]


@pytest.fixture(scope="module")
def built_lib():
    from ssf_slam_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    decls = _header_decls()
    assert len(decls) >= 25
    for name in decls:
        assert hasattr(lib, name), name
    assert lib.ssf_abi_version() == 1


def test_ctypes_table_matches_header(built_lib):
    from ssf_slam_b200 import _native
    decls = _header_decls()
    assert set(decls) == set(_native.SIGNATURES)
    for name, args in decls.items():
        assert "".join(_kind(a) for a in args) == _native.SIGNATURES[name][0], name


def test_dense_args_struct_layout(built_lib):
    """The ctypes mirror of ssf_dense_args has the size the compiled library sees."""
    from ssf_slam_b200 import _native
    lib = ctypes.CDLL(built_lib)
    assert lib.ssf_dense_args_bytes() == ctypes.sizeof(_native.DenseArgs)


def test_no_cpu_fallback():
    import torch
    from ssf_slam_b200 import _native, pointnet2_utils
    with pytest.raises(_native.SsfError):
        if torch.cuda.is_available():
            pointnet2_utils.knn(3, torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))  # CPU tensors are refused
        else:
            pointnet2_utils.furthest_point_sample(torch.zeros(1, 8, 3), 2)      # no device -> refuse, never compute on CPU
