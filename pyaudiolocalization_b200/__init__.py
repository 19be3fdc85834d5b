"""pyaudiolocalization_b200 -- B200-native (sm_100a) implementation of the data-parallel hot
path of zeynelacikgoez/PyAudioLocalization: batched multipath scene synthesis and batched
GCC-PHAT / TDOA estimation, behind the reference's own function signatures.

The compute lives in libpal_b200.so (hand-written CUDA, C ABI in include/pal_b200.h); this
package is the thin host layer.  There is no CPU fallback.
"""
from ._lib import PalError, launch_count  # noqa: F401
from . import filters, gcc_phat, shard, solver, synth  # noqa: F401
from .gcc_phat import TdoaBatch, all_pairs, gcc_phat_tdoa_batched, peak_distance, window_half_width  # noqa: F401

__version__ = "0.1.0"
