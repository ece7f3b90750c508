"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/freesplat_b200.h declares, and the
ctypes argument structs have exactly the compiled sizes (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from freesplat_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "freesplat_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from freesplat_b200 import build
        build.build()
    return C.CDLL(_lib.LIB_PATH)


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t|const char\*)\s+(fs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms, (sorted(set(syms) ^ set(_lib.EXPORTS)))


def test_abi_version_and_error_string(lib):
    assert lib.fs_abi_version() == _lib.ABI_VERSION
    lib.fs_last_error.restype = C.c_char_p
    assert isinstance(lib.fs_last_error(), bytes)


def test_struct_sizes_match_ctypes(lib):
    from freesplat_b200.adapter import FsAdapterArgs, FsBackprojectArgs
    from freesplat_b200.cost_volume import FsCostVolumeArgs
    from freesplat_b200.depth_head import FsDepthHeadArgs
    from freesplat_b200.ply_export import FsPlyArgs
    from freesplat_b200.ptf import FsGruBwdDataArgs, FsGruBwdWeightsArgs, FsPtfArgs, FsPtfGruArgs, FsPtfMergeBwdArgs
    for which, st in enumerate([_lib.FsRasterFwdArgs, _lib.FsRasterBwdArgs, FsCostVolumeArgs, FsPtfArgs, FsPtfGruArgs, FsAdapterArgs,
                               FsDepthHeadArgs, FsBackprojectArgs, FsPlyArgs, FsPtfMergeBwdArgs]):
        assert lib.fs_struct_size(which) == C.sizeof(st), (which, st.__name__, lib.fs_struct_size(which), C.sizeof(st))
    for which, st in ((12, FsGruBwdDataArgs), (13, FsGruBwdWeightsArgs)):
        assert lib.fs_struct_size(which) == C.sizeof(st), (which, st.__name__, lib.fs_struct_size(which), C.sizeof(st))
    assert lib.fs_struct_size(99) == -1


def test_invalid_arguments_are_rejected_without_a_gpu(lib):
    """Argument validation happens before any CUDA call: NULL / inconsistent arguments return FS_ERR_INVALID_ARG."""
    a = _lib.FsRasterFwdArgs(P=10, V=1, H=16, W=16)
    assert lib.fs_raster_forward(C.byref(a), None) == -1
    assert lib.fs_raster_forward(None, None) == -1
    lib.fs_last_error.restype = C.c_char_p
    assert b"invalid argument" in lib.fs_last_error()
