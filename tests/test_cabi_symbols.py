"""The built C-ABI library loads without a GPU and exports every symbol include/nemo_b200.h declares."""
import ctypes
import os

from nemo_cvpr2023_b200 import _lib, build


def test_library_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    protos = _lib.parse_header()
    assert len(protos) >= 40
    missing = [n for n in protos if not hasattr(dll, n)]
    assert not missing, missing
    assert dll.nemo_b200_version() >= 100


def test_binding_resolves_prototypes():
    _lib.lib.load()   # raises if the .so lacks a declared symbol
    assert 'nemo_b200_gemm_tc' in _lib.lib.protos and 'nemo_b200_skin_bwd' in _lib.lib.protos
