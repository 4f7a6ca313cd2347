"""ctypes binding of libb2s.so (include/b2s.h).  No torch types cross this boundary:
every pointer is a plain integer address (device pointers come from tensor.data_ptr()).

The product path has NO CPU fallback: load() raises if the CUDA library is missing and
Engine() raises if b2s_init() finds no sm_100 device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2s.so")

TPL_MAX_BYTES = 2048

ERR_ASSERT_NPO2, ERR_ASSERT_ROOT, ERR_ASSERT_PRIMITIVE = -10, -11, -12


class LeafTemplates(C.Structure):
    """struct b2s_leaf_templates"""
    _fields_ = [("n_slots", C.c_uint32), ("trim", C.c_uint32),
                ("seg_off", (C.c_uint32 * 5) * 4), ("bytes", C.c_uint8 * TPL_MAX_BYTES)]


class B2SError(RuntimeError):
    pass


# every symbol include/b2s.h declares: name -> (restype, argtypes)
_u64, _u32, _vp, _int = C.c_uint64, C.c_uint32, C.c_void_p, C.c_int
_TP = C.POINTER(LeafTemplates)
SIGNATURES = {
    "b2s_version": (_int, []),
    "b2s_last_error": (C.c_char_p, []),
    "b2s_init": (_int, [_int]),
    "b2s_shutdown": (_int, []),
    "b2s_trim": (_int, []),
    "b2s_launch_count": (_u64, []),
    "b2s_device_sm_count": (_int, []),
    "b2s_gl_mul": (_u64, [_u64, _u64]),
    "b2s_gl_pow": (_u64, [_u64, _u64]),
    "b2s_gl_inv": (_u64, [_u64]),
    "b2s_ntt": (_int, [_vp, _u64, _u32, _vp, _u64, _u32, _u32, _u64, _u64, _int, _vp]),
    "b2s_ntt_host": (_int, [_vp, _u64, _u32, _vp, _u64, _u32, _u32, _u64, _u64, _int]),
    "b2s_scale": (_int, [_vp, _u64, _vp, _u64, _u64, _u32, C.POINTER(_u64), _vp]),
    "b2s_eval_points": (_int, [_vp, _u64, _u32, _u64, _vp, _u64, _u32, _u64, _vp, _u64, _vp]),
    "b2s_merkle_field": (_int, [_vp, _u64, _u64, _TP, _vp, _vp]),
    "b2s_merkle_blobs": (_int, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "b2s_merkle_rows": (_int, [_vp, _vp, _u32, _u64, _vp, _vp, _u32, _vp, _u32, _vp, _u32, _vp, _u32, _vp, _u64, _vp,
                               _int, _vp, _vp, _vp]),
    "b2s_merkle_upper": (_int, [_vp, _u64, _vp]),
    "b2s_merkle_open": (_int, [_vp, _u64, _vp, _u32, _vp, _vp]),
    "b2s_fri_fold": (_int, [_vp, _u64, _u64, C.POINTER(_u64), _u64, _u64, _vp, _u64, _TP, _vp, _vp]),
    "b2s_gather": (_int, [_vp, _u64, _u32, _vp, _u32, _vp, _vp]),
    "b2s_quotients": (_int, [_vp, _u64, _u32, _u64, _u32, _vp, _vp, _vp, _u32, _u32, _u64, _u64, _u64, _u64, _vp,
                             C.POINTER(C.c_int), _vp, _vp]),
    "b2s_open_multi": (_int, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp]),
    "b2s_combination": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u64, _u64, _u64, _vp, _u64, _vp]),
    "b2s_dist_twiddle_transpose": (_int, [_vp, _u64, _u32, _u32, _u64, _u64, _u64, C.POINTER(_vp), _u32, _u64, _u64,
                                          _vp]),
    "b2s_block_permute": (_int, [_vp, _vp, _u32, _u32, _u32, _vp]),
    "b2s_ntt_timed": (_int, [_vp, _u64, _u32, _vp, _u64, _u32, _u32, _u64, _u64, _int, _vp, _u32,
                             C.POINTER(C.c_float)]),
}

_lib = None


def load(path=None):
    """dlopen libb2s.so and attach the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise B2SError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib, rc):
    """status code -> exception.  The reference's `assert`s surface as AssertionError
    (code/ntt.py:5-6, :13-16, :29-36)."""
    if rc == 0:
        return
    msg = lib.b2s_last_error()
    msg = msg.decode() if isinstance(msg, bytes) else str(msg)
    if rc in (ERR_ASSERT_NPO2, ERR_ASSERT_ROOT, ERR_ASSERT_PRIMITIVE):
        raise AssertionError(msg)
    raise B2SError("libb2s error %d: %s" % (rc, msg))
