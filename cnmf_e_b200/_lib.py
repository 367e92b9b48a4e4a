"""ctypes loader for libcnmfe_b200.so (the C ABI declared in include/cnmfe_b200.h).

The product has no CPU fallback: if the library is missing or has no CUDA device, calls raise."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNMFE_B200_LIB") or os.path.join(_HERE, "libcnmfe_b200.so")   # env override: development A/B builds only

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_float_p = ctypes.POINTER(ctypes.c_float)


class DeconvOpts(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("method", ctypes.c_int), ("optimize_b", ctypes.c_int),
                ("optimize_pars", ctypes.c_int), ("maxIter", ctypes.c_int), ("has_tau_range", ctypes.c_int),
                ("smin", ctypes.c_double), ("lam", ctypes.c_double), ("b", ctypes.c_double),
                ("max_tau", ctypes.c_double), ("tau_range", ctypes.c_double * 2),
                ("thresh_factor", ctypes.c_double), ("p_noise", ctypes.c_double)]


class Options(ctypes.Structure):
    _fields_ = [("spatial_algorithm", ctypes.c_int), ("maxIter_temporal", ctypes.c_int),
                ("deconv_flag", ctypes.c_int), ("bg_acceleration", ctypes.c_int),
                ("replicate_spatial_aprev_quirk", ctypes.c_int), ("use_tensor_gram", ctypes.c_int),
                ("deconv", DeconvOpts), ("background_model", ctypes.c_int), ("nb", ctypes.c_int), ("bg_ssub", ctypes.c_int),
                ("thresh_outlier", ctypes.c_double)]


class CnmfeError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/cnmfe_b200.h declares
V = ctypes.c_void_p
I = ctypes.c_int
SYMBOLS = {
    "cnmfe_last_error": (ctypes.c_char_p, []),
    "cnmfe_deconv_defaults": (None, [ctypes.POINTER(DeconvOpts)]),
    "cnmfe_options_defaults": (None, [ctypes.POINTER(Options)]),
    "cnmfe_launch_count": (ctypes.c_ulonglong, []),
    "cnmfe_deconvolve": (I, [V, I, I, ctypes.POINTER(DeconvOpts), V, V, V, V, V, V, V, V, V, I]),
    "cnmfe_deconvolve_dev": (I, [V, I, I, ctypes.POINTER(DeconvOpts), V, V, V, V, V, I]),
    "cnmfe_get_sn": (I, [V, I, I, V, I]),
    "cnmfe_hals_temporal_uv": (I, [V, V, I, I, V, I, ctypes.POINTER(DeconvOpts), V, V, V, V, I]),
    "cnmfe_update_temporal_finish_part": (I, [V, I, I]),
    "cnmfe_temporal_state_buffers": (I, [V, V, V, V, V]),
    "cnmfe_graph_conn_comp": (I, [I, V, V, V, V]),
    "cnmfe_debug_local_view": (I, [I, I, V, V, I, V, V, V, I, I, V, V, V, V, V, V, V, V, V, V, V, V, V, V]),
    "cnmfe_connectivity_constraint": (I, [I, I, I, V, V, V, ctypes.c_double, I]),
    "cnmfe_circular_constraints": (I, [I, I, I, V, V, V, V, V, V, ctypes.c_int64]),
    "cnmfe_search_location_dilate": (I, [I, I, I, V, V, V, ctypes.c_double, I, I, V, V, ctypes.c_int64]),
    "cnmfe_search_location_ellipse": (I, [I, I, I, V, V, V, ctypes.c_double, ctypes.c_double, ctypes.c_double, V, V, ctypes.c_int64]),
    "cnmfe_create": (I, [ctypes.POINTER(V), I, I, I, I, V, V, V, I, I, I]),
    "cnmfe_destroy": (None, [V]),
    "cnmfe_set_options": (I, [V, ctypes.POINTER(Options)]),
    "cnmfe_upload_block": (I, [V, I, V, I]),
    "cnmfe_upload_block_dev": (I, [V, I, V, I]),
    "cnmfe_set_neurons": (I, [V, I, V, V, V, V]),
    "cnmfe_set_prev": (I, [V, I, V, V, V, V]),
    "cnmfe_set_search": (I, [V, I, V, V]),
    "cnmfe_set_sn": (I, [V, V]),
    "cnmfe_ring_offsets": (I, [V, c_int_p, V, V]),
    "cnmfe_set_ring": (I, [V, I, V, V]),
    "cnmfe_get_ring": (I, [V, I, V, V]),
    "cnmfe_ssub_dims": (I, [V, I, c_int_p, c_int_p, c_int_p, V, V]),
    "cnmfe_set_bf": (I, [V, I, V, V, V]),
    "cnmfe_get_bf": (I, [V, I, V, V, V]),
    "cnmfe_update_background": (I, [V]),
    "cnmfe_update_spatial": (I, [V]),
    "cnmfe_update_spatial_ex": (I, [V, I]),
    "cnmfe_get_sn_map": (I, [V, V]),
    "cnmfe_get_spatial": (I, [V, V]),
    "cnmfe_set_spatial": (I, [V, V]),
    "cnmfe_update_temporal_patches": (I, [V]),
    "cnmfe_temporal_merge_buffers": (I, [V, ctypes.POINTER(V), ctypes.POINTER(V)]),
    "cnmfe_update_temporal_finish": (I, [V]),
    "cnmfe_update_temporal": (I, [V]),
    "cnmfe_set_use_c_hat": (I, [V, I]),
    "cnmfe_set_trace_major": (I, [V, I]),
    "cnmfe_host_register": (I, [V, ctypes.c_size_t]),
    "cnmfe_host_unregister": (I, [V]),
    "cnmfe_get_temporal": (I, [V, V, V, V, V, V]),
    "cnmfe_compute_rss": (I, [V, I, I, V, V, V]),
    "cnmfe_reconstruct_background": (I, [V, I, I, I, V, V, V]),
    "cnmfe_sync": (I, [V]),
    "cnmfe_timer_begin": (I, [V]),
    "cnmfe_timer_end": (I, [V, c_float_p]),
    "cnmfe_last_phase_ms": (I, [V, c_float_p]),
    "cnmfe_debug_second_moments": (I, [V, I, I, V]),
    "cnmfe_last_gram_was_tensor": (I, [V]),
    "cnmfe_last_gram_frames": (I, [V]),
    "cnmfe_last_active_pixels": (ctypes.c_longlong, [V]),
    "cnmfe_last_nmf_iterations": (I, [V]),
    "cnmfe_debug_video_rows": (I, [V, I, I, V, V]),
    "cnmfe_get_merged_craw": (I, [V, V]),
    "cnmfe_estimate_noise": (I, [V, I, I, V]),
}


def lib():
    """Load the shared library (raises if it was not built: there is no fallback implementation)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CnmfeError("libcnmfe_b200.so not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                             "cnmf_e_b200 has no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        missing = [n for n in SYMBOLS if not hasattr(L, n)]
        if missing:
            raise CnmfeError("libcnmfe_b200.so is stale: missing symbols %s (rebuild)" % missing)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise CnmfeError(lib().cnmfe_last_error().decode("utf-8", "replace"))
