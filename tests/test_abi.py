"""CPU: the C-ABI library loads and exports every symbol include/dpc_b200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'dpc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dpc_[a-z0-9_]+)\s*\(', src)))


def test_header_matches_binding_table():
    from dpc_b200._lib import EXPORTS
    assert sorted(EXPORTS) == header_functions()


def test_library_loads_and_exports_every_symbol():
    from dpc_b200._lib import LIB_PATH, lib
    assert os.path.exists(LIB_PATH), 'build the library first: python -m dpc_b200.build'
    dll = ctypes.CDLL(LIB_PATH)
    for name in header_functions():
        assert hasattr(dll, name), name
    L = lib()
    assert L.abi_version() == 1
    assert L.launch_count() >= 0


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    from dpc_b200._lib import _Lib, DpcLibError
    with pytest.raises(DpcLibError):
        _Lib(str(tmp_path / 'nope.so'))


def test_bad_arguments_return_error_codes_not_crashes():
    """argument validation happens before any CUDA call, so it is testable without a GPU"""
    import pytest
    from dpc_b200._lib import lib, DpcLibError, ConvGeom
    L = lib()
    with pytest.raises(DpcLibError, match='bad dims'):
        L.gemm_f32(0, 0, 0, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, None)
    g = ConvGeom(1, 1, 4, 4, 3, 1, 4, 4, 64, 1, 3, 3, 1, 1, 1, 0, 1, 1)       # Ci = 3: unsupported
    with pytest.raises(DpcLibError, match='multiples of 64'):
        L.conv3d_fwd_tc(g, 1, 1, 1, 1, 1, None, None)
    g = ConvGeom(1, 1, 4, 4, 64, 1, 5, 4, 64, 1, 3, 3, 1, 1, 1, 0, 1, 1)       # wrong Ho
    with pytest.raises(DpcLibError, match='output extent'):
        L.conv3d_fwd_tc(g, 1, 1, 1, 1, 1, None, None)
    with pytest.raises(DpcLibError, match='unsupported channel count'):
        L.bn_stats(1, 10, 6, 1, 1, 1, 1e-5, None)
