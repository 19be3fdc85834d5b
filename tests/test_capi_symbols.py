"""The C-ABI library loads without a GPU and exports every symbol include/pal_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pal_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_declared_symbols():
    from pyaudiolocalization_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert {"pal_abi_version", "pal_last_error", "pal_gcc_phat_tdoa", "pal_gcc_phat_workspace",
            "pal_profile_hook", "pal_launch_count"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pal_abi_version() == 8


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so it can be exercised on the CPU box."""
    from pyaudiolocalization_b200 import _lib
    L = _lib.lib()
    full, small = ctypes.c_size_t(), ctypes.c_size_t()
    assert L.pal_gcc_phat_workspace(16, 4, 2048, 6, ctypes.byref(full), ctypes.byref(small)) == 0
    assert full.value >= 16 * 4 * 2080 * 8 and small.value < full.value
    assert L.pal_gcc_phat_workspace(16, 1, 2048, 6, ctypes.byref(full), None) == -1
    prm = _lib.TdoaParams(800, 0, 0, 1.0, 1, 2e-6, 1, 0, 0)
    rc = L.pal_gcc_phat_tdoa(None, 1, 4, 2048, None, 6, ctypes.byref(prm), None, None, None, None, None, None,
                             None, 0, None)
    assert rc == -1 and b"NULL" in L.pal_last_error()


def test_host_side_integer_decisions():
    import pyaudiolocalization_b200 as pal
    from oracle import pal_oracle as O
    for fs in (8000.0, 16000.0, 44100.0, 48000.0):
        for med in (None, 0.05, 0.01, 0.0005, 1.0, -1.0, 0.0):
            assert pal.window_half_width(2048, 2048, fs, med) == O.window_half_width(2048, 2048, fs, med)
        assert pal.peak_distance(fs) == O.peak_distance(fs)
    assert [tuple(x) for x in pal.all_pairs(3)] == [(0, 1), (0, 2), (1, 2)]
