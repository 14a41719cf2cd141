"""ctypes binding of libendosurf_b200.so (the C ABI declared in include/endosurf_b200.h).

There is deliberately NO fallback: if the shared library is missing or the device is not an sm_100 GPU the
import / context creation raises.  The product path never routes through PyTorch ops or the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ES_LIB_PATH: profiling tools load an instrumented variant of the library (endosurf_b200.build --tag=...); the product
# never sets it
LIB_PATH = os.environ.get("ES_LIB_PATH") or os.path.join(_HERE, "libendosurf_b200.so")

ES_E = {-1: "ES_E_BADARG", -2: "ES_E_UNSUPPORTED", -3: "ES_E_NOWEIGHTS", -4: "ES_E_DEVICE"}


class EsNetConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "use_deform", "n_layers", "skip_layer", "hidden_dim", "multires_deform_pos", "multires_deform_time",
        "multires_sdf_pos", "multires_color_pos", "multires_color_dir", "precision_terms")]


class EsRenderParams(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_importance", C.c_int32), ("up_sample_steps", C.c_int32),
                ("do_upsample", C.c_int32), ("cos_anneal_ratio", C.c_float), ("variance", C.c_void_p),
                ("t_vals", C.c_void_p), ("u_vals", C.c_void_p), ("t_rand", C.c_void_p), ("z_override", C.c_void_p)]


class EsRenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "color_map", "depth_map", "gradients_o", "gradient_o_error", "weights", "weight_max", "cdf", "s_val",
        "z_vals", "sdf", "sampled_color")]


class EsProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 12), ("launches", C.c_int64 * 12), ("points", C.c_int64 * 12)]


class EsTrainParams(C.Structure):
    """es_train_params: [net] -> table of n_layers device pointers."""
    _fields_ = [(n, C.POINTER(C.c_void_p) * 3) for n in ("v", "g", "grad_v", "grad_g", "grad_b")]


class EsRenderGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "color_map", "depth_map", "gradients_o", "gradient_o_error", "weights", "cdf", "sdf", "sampled_color")]


EXPORTS = {
    # name: (restype, argtypes)
    "es_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(EsNetConfig)]),
    "es_destroy": (None, [C.c_void_p]),
    "es_last_error": (C.c_char_p, [C.c_void_p]),
    "es_sync_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "es_poll_error": (C.c_int, [C.c_void_p, C.c_void_p]),
    "es_num_sms": (C.c_int, [C.c_void_p]),
    "es_release_workspace": (C.c_int, [C.c_void_p]),
    "es_load_network": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "es_sdf_query": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                               C.c_void_p]),
    "es_sdf_grid": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "es_point_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "es_load_network_wn": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.c_void_p]),
    "es_set_plane_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "es_train_stash_bytes": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "es_point_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                         C.c_int64, C.c_int64, C.c_int64] + [C.c_void_p] * 7),
    "es_point_train_backward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64] +
                                [C.c_void_p] * 9 + [C.POINTER(EsTrainParams), C.c_void_p]),
    "es_render_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                          C.c_float] + [C.c_void_p] * 7 + [C.POINTER(EsRenderOut), C.c_void_p,
                                                                           C.c_void_p]),
    "es_render_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_float] + [C.c_void_p] * 8 +
                                 [C.POINTER(EsRenderGrads), C.POINTER(EsTrainParams), C.c_void_p, C.c_void_p]),
    "es_debug_set": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "es_wgrad_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "es_up_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "es_render_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(EsRenderParams),
                                 C.POINTER(EsRenderOut), C.c_void_p]),
    "es_launch_count": (C.c_int64, [C.c_void_p]),
    "es_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "es_profile_read": (C.c_int, [C.c_void_p, C.POINTER(EsProfile), C.c_void_p]),
    "es_debug_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "es_mma_bench": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int64)]),
    "es_chunk_colmap": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "es_umma_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the library and declare every prototype; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m endosurf_b200.build` (nvcc, sm_100a). "
            "endosurf_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EsError(RuntimeError):
    pass


def check(ctx, rc: int, what: str):
    if rc == 0:
        return
    lib = load()
    msg = lib.es_last_error(ctx).decode() if ctx else ""
    code = ES_E.get(rc, f"cudaError {rc}")
    raise EsError(f"{what} failed: {code} {msg}")
