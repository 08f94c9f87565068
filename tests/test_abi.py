"""The C-ABI shared library loads and exports every symbol include/*.h declares."""

import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(hdr).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(swiftly_b200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_eight_primitives():
    syms = declared_symbols()
    for prim in ("prepare_facet", "extract_from_facet", "add_to_subgrid", "finish_subgrid",
                 "prepare_subgrid", "extract_from_subgrid", "add_to_facet", "finish_facet"):
        assert f"swiftly_b200_{prim}" in syms


def test_library_exports_every_declared_symbol():
    from ska_sdp_distributed_fourier_transform_b200 import _lib, build

    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # the binding table covers the header exactly
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    lib.swiftly_b200_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.swiftly_b200_build_info()


def test_missing_library_fails_loudly(tmp_path):
    from ska_sdp_distributed_fourier_transform_b200 import _lib

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "libswiftly_b200.so"))
