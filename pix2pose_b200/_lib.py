"""ctypes binding of libpix2pose_b200.so (include/pix2pose_b200.h).  Fails loudly: no CPU fallback."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2P_LIB", os.path.join(_HERE, "libpix2pose_b200.so"))   # P2P_LIB: A/B experiments with another build

P2P_OK = 0
PREC_FP16X3 = 0
PREC_FP16 = 1
_PRECISIONS = {"fp16x3": PREC_FP16X3, "fp16": PREC_FP16}

_lib = None


class P2PError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pix2pose_b200 error %d: %s" % (code, msg))
        self.code = code


def precision_code(p):
    if isinstance(p, str):
        if p not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        return _PRECISIONS[p]
    return int(p)


def lib():
    """Loads the shared library (building it is `__graft_entry__.build()` / csrc/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m pix2pose_b200.csrc.build` "
                          "(nvcc, sm_100a). pix2pose_b200 has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_f = ctypes.POINTER(ctypes.c_float)
    c_d = ctypes.POINTER(ctypes.c_double)
    c_i = ctypes.POINTER(ctypes.c_int)
    vp = ctypes.c_void_p
    sig = {
        "p2p_last_error": (ctypes.c_char_p, []),
        "p2p_version": (ctypes.c_char_p, []),
        "p2p_set_device": (ctypes.c_int, [ctypes.c_int]),
        "p2p_param_count": (ctypes.c_size_t, [ctypes.c_char_p]),
        "p2p_flops_per_crop": (ctypes.c_double, [ctypes.c_char_p]),
        "p2p_engine_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]),
        "p2p_engine_destroy": (None, [vp]),
        "p2p_engine_capacity": (ctypes.c_int, [vp]),
        "p2p_engine_launch_count": (ctypes.c_longlong, [vp]),
        "p2p_model_create": (ctypes.c_int, [vp, c_f, ctypes.c_size_t, ctypes.POINTER(vp)]),
        "p2p_model_destroy": (None, [vp]),
        "p2p_predict": (ctypes.c_int, [vp, vp, c_f, ctypes.c_int, c_f, c_f]),
        "p2p_predict_device": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, vp, vp, vp]),
        "p2p_pnp_ransac": (ctypes.c_int, [c_d, c_d, ctypes.c_int, c_d, ctypes.c_float, ctypes.c_int, ctypes.c_double,
                                          c_d, c_d, c_d, c_i, ctypes.POINTER(ctypes.c_uint8), c_i]),
        "p2p_pnp_ransac_batch": (ctypes.c_int, [c_d, c_d, c_i, ctypes.c_int, c_d, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                                ctypes.c_double, vp, ctypes.POINTER(ctypes.c_uint8), c_f]),
        "p2p_pipeline_create": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]),
        "p2p_pipeline_destroy": (None, [vp]),
        "p2p_pipeline_run": (ctypes.c_int, [vp, vp, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int, vp,
                                            ctypes.c_int, c_d, ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_double, vp]),
        "p2p_pipeline_run_f32": (ctypes.c_int, [vp, vp, c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp,
                                                ctypes.c_int, c_d, ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_double, vp]),
        "p2p_pipeline_run_multi": (ctypes.c_int, [vp, ctypes.POINTER(vp), c_i, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, vp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_double, vp]),
        "p2p_pipeline_run_device": (ctypes.c_int, [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp,
                                                   ctypes.c_int, c_d, ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_double, vp]),
        "p2p_pipeline_fetch_crop": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint8)]),
        "p2p_pipeline_fetch_decode": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_f]),
        "p2p_pipeline_fetch_buffer": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_f]),
        "p2p_depth_xyz": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]),
        "p2p_depth_normals": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_int), ctypes.c_double,
                                             ctypes.POINTER(ctypes.c_double)]),
        "p2p_pipeline_mask_iou": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_longlong)]),
        "p2p_pipeline_set_box_size": (ctypes.c_int, [vp, ctypes.c_double]),
        "p2p_pipeline_debug_override": (ctypes.c_int, [vp, ctypes.c_int, c_f, c_f, ctypes.c_int]),
        "p2p_pipeline_launch_count": (ctypes.c_longlong, [vp]),
        "p2p_pipeline_forward_ms": (ctypes.c_int, [vp, c_d]),
        "p2p_pipeline_set_async": (ctypes.c_int, [vp, ctypes.c_int]),
        "p2p_pipeline_wait": (ctypes.c_int, [vp, vp]),
        "p2p_pipeline_debug_select": (ctypes.c_int, [vp, c_d, c_i, c_i, ctypes.c_int, ctypes.c_int, vp]),
        "p2p_time_forward": (ctypes.c_int, [vp, vp, c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f]),
        "p2p_engine_event_record": (ctypes.c_int, [vp, ctypes.c_int]),
        "p2p_engine_event_elapsed": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_f]),
        "p2p_engine_prof_begin": (ctypes.c_int, [vp]),
        "p2p_engine_prof_end": (ctypes.c_int, [vp, c_d, c_i]),
        "p2p_engine_profile_forward": (ctypes.c_int, [vp, vp, c_f, ctypes.c_int, c_d, c_i]),
        "p2p_pipeline_upload_frames": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                      ctypes.POINTER(vp)]),
        "p2p_host_alloc": (vp, [ctypes.c_size_t]),
        "p2p_host_free": (None, [vp]),
        "p2p_engine_read_tensor": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_int, c_f, c_i, c_i, c_i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._sig_names = sorted(sig)
    _lib = L
    return L


def check(code):
    if code != P2P_OK:
        raise P2PError(code, lib().p2p_last_error().decode("utf-8", "replace"))


def fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def pinned_array(shape, dtype=np.float32):
    """numpy view over cudaMallocHost memory (freed when the returned array's base is collected)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().p2p_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("cudaMallocHost(%d) failed" % n)
    buf = (ctypes.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _pinned_keep[arr.ctypes.data] = p
    return arr


_pinned_keep = {}
