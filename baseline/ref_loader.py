"""Import the unmodified reference from baseline/_ref (see tools/install_reference.py) for the CPU arm of bench.py.

soundfile, resampy and matplotlib are not installed in this image and are never touched by the hot path
(`sf` only in utils.py:469, `resampy` only in signal_processing.py:106, `plt` only in plotting code), so empty stub
modules stand in for them; nothing of the reference itself is altered."""
import importlib
import os
import sys
import types

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "utils.py"))


def load():
    """-> the reference's `utils` module (phat_correlation, get_time_delays_phat, ...)."""
    for name in ("soundfile", "resampy", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:          # noqa: BLE001 - absent package: stub it
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["mpl_toolkits.mplot3d"], "Axes3D"):
        sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return importlib.import_module("utils")
