"""ctypes binding of libfocus_savi.so (the C ABI declared in include/focus_savi.h).

The product path has no fallback: if the library is missing this module raises,
and every non-zero return code becomes a RuntimeError carrying savi_last_error().
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FOCUS_SAVI_LIB") or os.path.join(_HERE, "libfocus_savi.so")     # override: development A/B builds

SAVI_DTYPE_F32 = 0
SAVI_DTYPE_BF16 = 1
SAVI_MAX_BLOCKS = 4
SAVI_MAX_SLOTS = 64
PATH_NAMES = {0: "simt-fp32", 1: "mma.sync-bf16", 2: "tcgen05-bf16"}

EXPORTS = ["savi_version", "savi_last_error", "savi_query", "savi_param_layout", "savi_pack_params",
           "savi_forward", "savi_backward", "savi_last_launch_count", "savi_profile_enable", "savi_profile_read", "savi_debug_set_phase_buffer", "savi_set_option", "savi_allreduce_peers"]
EXPORTS_STEVE = ["steve_attention_overlay", "steve_ari_tables", "steve_token_mlp", "steve_token_mlp_ws_bytes"]            # include/focus_steve.h


class SaviShape(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "T", "N", "D", "Ds", "M", "K", "I", "blocks", "heads", "dtype", "cluster")] + \
               [("eps", ctypes.c_float), ("ln_eps", ctypes.c_float)]


class SaviSizes(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in
                ("n_params", "param_floats", "packed_bytes", "saved_bytes", "fwd_ws_bytes", "bwd_ws_bytes")] + \
               [("cluster", ctypes.c_int32), ("path", ctypes.c_int32), ("dropout_floats", ctypes.c_int64)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "focus_b200: %s is missing. Build it with `python -m focus_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, ip, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.savi_version.restype = ip
    lib.savi_last_error.restype = ctypes.c_char_p
    lib.savi_last_launch_count.restype = ip
    lib.savi_query.argtypes = [ctypes.POINTER(SaviShape), ctypes.POINTER(SaviSizes)]
    lib.savi_query.restype = ip
    lib.savi_param_layout.argtypes = [ctypes.POINTER(SaviShape), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.savi_param_layout.restype = ip
    lib.savi_pack_params.argtypes = [ctypes.POINTER(SaviShape), ctypes.POINTER(vp), vp, vp]
    lib.savi_pack_params.restype = ip
    lib.savi_forward.argtypes = [ctypes.POINTER(SaviShape)] + [vp] * 9
    lib.savi_forward.restype = ip
    lib.savi_backward.argtypes = [ctypes.POINTER(SaviShape)] + [vp] * 12
    lib.savi_backward.restype = ip
    lib.savi_profile_enable.argtypes = [ip]
    lib.savi_profile_enable.restype = ip
    lib.savi_profile_read.argtypes = [ctypes.POINTER(ctypes.c_float), ip]
    lib.savi_profile_read.restype = ip
    lib.savi_debug_set_phase_buffer.argtypes = [vp]
    lib.savi_debug_set_phase_buffer.restype = ip
    lib.savi_set_option.argtypes = [ctypes.c_char_p, ip]
    lib.savi_set_option.restype = ip
    lib.savi_allreduce_peers.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), vp, ip, ip, vp, i64, ctypes.c_float, i64, vp]
    lib.savi_allreduce_peers.restype = ip
    lib.steve_attention_overlay.argtypes = [vp, ip, vp, vp, vp, i64, ip, ip, ip, ip, ip, ip, vp]
    lib.steve_attention_overlay.restype = ip
    lib.steve_ari_tables.argtypes = [vp, vp, vp, ip, ip, ip, i64, vp]
    lib.steve_ari_tables.restype = ip
    lib.steve_token_mlp_ws_bytes.argtypes = [ip]
    lib.steve_token_mlp_ws_bytes.restype = i64
    lib.steve_token_mlp.argtypes = [vp] * 8 + [ip, i64, ip, ip, ctypes.c_float, vp, vp]
    lib.steve_token_mlp.restype = ip
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        raise RuntimeError("focus_b200 %s failed (code %d): %s" % (what, rc, lib.savi_last_error().decode()))


def set_option(name, value):
    """Process-wide development option of the library (include/focus_savi.h: savi_set_option)."""
    check(lib.savi_set_option(name.encode(), int(value)), "savi_set_option")


def query(shape):
    sizes = SaviSizes()
    check(lib.savi_query(ctypes.byref(shape), ctypes.byref(sizes)), "savi_query")
    return sizes


def param_layout(shape, n_params):
    off = (ctypes.c_int64 * n_params)()
    num = (ctypes.c_int64 * n_params)()
    check(lib.savi_param_layout(ctypes.byref(shape), off, num), "savi_param_layout")
    return list(off), list(num)


PROFILE_SLOTS = ["pack_params", "ln_tokens_fwd", "savi_fwd", "savi_bwd", "wgrad", "ln_tokens_bwd"]


def profile_enable(on=True):
    check(lib.savi_profile_enable(1 if on else 0), "savi_profile_enable")


def profile_read():
    """{kernel: ms} of the last forward + backward (device-synchronising)."""
    buf = (ctypes.c_float * len(PROFILE_SLOTS))()
    check(lib.savi_profile_read(buf, len(PROFILE_SLOTS)), "savi_profile_read")
    return {k: float(v) for k, v in zip(PROFILE_SLOTS, buf)}
