"""ctypes binding of libpal_b200.so (the C ABI declared in include/pal_b200.h).

There is no CPU fallback: if the CUDA library has not been built, importing any compute
entry point raises.  Build it with `python -m pyaudiolocalization_b200.build`.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PAL_B200_LIB") or os.path.join(_PKG, "libpal_b200.so")   # env override: tuning experiments only

PAL_ABI_VERSION = 8

# per-row flag bits (include/pal_b200.h)
FLAG_NEAR_TIE = 1
FLAG_CHAIN = 2
FLAG_PLATEAU = 4
FLAG_REFINED = 8
FLAG_FALLBACK_ARGMAX = 16
FLAG_ALT_THRESHOLD = 32
FLAG_STACK_OVERFLOW = 64


class PalError(RuntimeError):
    pass


class TdoaParams(C.Structure):
    _fields_ = [("win_half", C.c_int32), ("peak_dist", C.c_int32), ("thr_method", C.c_int32),
                ("thr_mult", C.c_float), ("num_peaks", C.c_int32), ("tie_eps", C.c_float),
                ("refine", C.c_int32), ("len_first", C.c_int32), ("len_second", C.c_int32)]


_lib = None


def lib():
    """Load the shared library once; fail loudly when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PalError(f"{LIB_PATH} not found: the CUDA extension is not built "
                       "(run `python -m pyaudiolocalization_b200.build`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.pal_abi_version.restype = C.c_int
    L.pal_last_error.restype = C.c_char_p
    L.pal_launch_count.restype = C.c_ulonglong
    L.pal_profile_hook.restype = C.c_int
    L.pal_profile_hook.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
    L.pal_gcc_phat_workspace.restype = C.c_int
    L.pal_gcc_phat_workspace.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.pal_gcc_phat_tdoa.restype = C.c_int
    L.pal_gcc_phat_tdoa.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                    C.POINTER(TdoaParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.pal_gcc_phat_tdoa_f64.restype = C.c_int
    L.pal_gcc_phat_tdoa_f64.argtypes = L.pal_gcc_phat_tdoa.argtypes
    L.pal_tdoa_seconds.restype = C.c_int
    L.pal_tdoa_seconds.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
    VP, I32, I64, F64, F32, SZP = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float, C.POINTER(C.c_size_t)
    L.pal_image_sources_workspace.restype = C.c_int
    L.pal_image_sources_workspace.argtypes = [I32, I32, I64, SZP]
    L.pal_image_sources.restype = C.c_int
    L.pal_image_sources.argtypes = [VP, I64, VP, I64, VP, I32, VP, VP, VP, I32, I64, I32, F64, F64, I32, I32, VP, VP, VP,
                                    VP, C.c_size_t, VP]
    L.pal_path_table_batched.restype = C.c_int
    L.pal_path_table_batched.argtypes = [VP, VP, VP, VP, I64, I32, VP, I32, I64, VP, VP, I32, F64, F64, I32, VP, VP, VP, VP, VP]
    L.pal_render_scenes_workspace.restype = C.c_int
    L.pal_render_scenes_workspace.argtypes = [I32, I64, SZP, SZP]
    L.pal_render_scenes.restype = C.c_int
    L.pal_render_scenes.argtypes = [VP, I32, I32, VP, VP, VP, I32, VP, I64, I32, F64, I32, VP, VP, C.c_size_t, VP]
    L.pal_render_plan_bytes.restype = C.c_int
    L.pal_render_plan_bytes.argtypes = [I32, SZP, SZP]
    L.pal_render_plan.restype = C.c_int
    L.pal_render_plan.argtypes = [VP, I32, I32, VP, C.c_size_t, VP, C.c_size_t, VP]
    L.pal_render_rows_workspace.restype = C.c_int
    L.pal_render_rows_workspace.argtypes = [I32, I64, SZP, SZP]
    L.pal_render_scenes_planned.restype = C.c_int
    L.pal_render_scenes_planned.argtypes = [VP, C.c_size_t, I32, I32, VP, VP, VP, I32, VP, I64, I32, F64, I32, VP, VP, C.c_size_t, VP]
    L.pal_path_table.restype = C.c_int
    L.pal_path_table.argtypes = [VP, VP, VP, I32, VP, I32, VP, VP, I32, F64, F64, VP, VP, VP]
    L.pal_render_workspace.restype = C.c_int
    L.pal_render_workspace.argtypes = [I32, I32, SZP, SZP]
    L.pal_render_scene.restype = C.c_int
    L.pal_render_scene.argtypes = [VP, I32, I32, VP, VP, I32, I32, F64, I32, I32, VP, VP, C.c_size_t, VP]
    L.pal_normalise_compress.restype = C.c_int
    L.pal_normalise_compress.argtypes = [VP, I64, I32, F32, F32, I32, VP]
    L.pal_filtfilt_workspace.restype = C.c_int
    L.pal_filtfilt_workspace.argtypes = [I64, I32, I32, SZP]
    L.pal_filtfilt.restype = C.c_int
    L.pal_filtfilt.argtypes = [VP, I64, I32, I32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), I32, I32,
                               VP, VP, C.c_size_t, VP]
    L.pal_pcm16_to_f32.restype = C.c_int
    L.pal_pcm16_to_f32.argtypes = [VP, I64, F32, VP, VP]
    L.pal_sync_align_workspace.restype = C.c_int
    L.pal_sync_align_workspace.argtypes = [I64, I32, I32, SZP, SZP]
    L.pal_sync_align.restype = C.c_int
    L.pal_sync_align.argtypes = [VP, I64, I32, I32, VP, VP, VP, VP, VP, VP, VP, C.c_size_t, VP]
    L.pal_pad_rows.restype = C.c_int
    L.pal_pad_rows.argtypes = [VP, I64, I64, VP, VP, VP, I64, I32, VP]
    L.pal_render_scenes_grouped.restype = C.c_int
    L.pal_render_scenes_grouped.argtypes = [I32, VP, VP, VP, VP, VP, VP, I32, VP, I32, F64, I32, VP, VP, C.c_size_t, VP]
    L.pal_reserve_sms.restype = C.c_int
    L.pal_reserve_sms.argtypes = [I32]
    L.pal_solve_positions_workspace.restype = C.c_int
    L.pal_solve_positions_workspace.argtypes = [I32, SZP]
    L.pal_solve_positions.restype = C.c_int
    L.pal_solve_positions.argtypes = [VP, I64, I32, VP, I32, VP, VP, VP, VP, VP, I64, F64, F64, I32, F64, F64, F64, VP, VP, VP, VP,
                                      C.c_size_t, VP]
    if L.pal_abi_version() != PAL_ABI_VERSION:
        raise PalError(f"libpal_b200.so ABI {L.pal_abi_version()} != expected {PAL_ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().pal_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise PalError(f"{what} failed ({rc}): {msg}")


def reserve_sms(n: int) -> None:
    """Leave `n` SMs out of the persistent grids of every later call (pal_reserve_sms), e.g. for a concurrent collective."""
    check(lib().pal_reserve_sms(int(n)), "pal_reserve_sms")


def launch_count() -> int:
    return int(lib().pal_launch_count())


def profile_hook(stage: int, start_event=None, stop_event=None):
    """Time one pipeline stage with the caller's CUDA events (see pal_profile_hook).  Pass
    torch.cuda.Event(enable_timing=True) objects that have been recorded once (so that their
    handles exist); stage 0 switches the hook off."""
    s = None if start_event is None else start_event.cuda_event
    e = None if stop_event is None else stop_event.cuda_event
    check(lib().pal_profile_hook(int(stage), s, e), "pal_profile_hook")
