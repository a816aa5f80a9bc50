"""ctypes binding of libb200canny.so (the C ABI declared in include/b200canny.h).

There is no fallback: if the CUDA library has not been built, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2C_LIB_PATH") or os.path.join(_HERE, "libb200canny.so")   # (override: profiling builds of the same library)

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_SIZE, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4, -5, -6
STAGE_MONO, STAGE_GAUSSIAN, STAGE_GRADIENT, STAGE_NMS, STAGE_THRESH, STAGE_HYSTER = range(6)
BUF_MONO, BUF_BLUR, BUF_GRAD, BUF_NMS, BUF_THRESH, BUF_EDGES, BUF_MAP2, BUF_BITS, BUF_VIEW = range(9)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C cudacam_b200/csrc` (or __graft_entry__.build()); "
        "there is no CPU fallback for the Canny path")

lib = C.CDLL(LIB_PATH)

_vp, _sz, _i, _u8p = C.c_void_p, C.c_size_t, C.c_int, C.c_void_p
_SIGS = {
    "b2c_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "b2c_destroy": (None, [_vp]),
    "b2c_set_low_threshold": (_i, [_vp, C.c_uint8]),
    "b2c_set_high_threshold": (_i, [_vp, C.c_uint8]),
    "b2c_get_low_threshold": (_i, [_vp]),
    "b2c_get_high_threshold": (_i, [_vp]),
    "b2c_enable_profiling": (_i, [_vp, _i]),
    "b2c_is_profiling_enabled": (_i, [_vp]),
    "b2c_last_timings": (_i, [_vp, C.POINTER(C.c_float), _i]),
    "b2c_run": (_i, [_vp, _u8p, _sz, _i]),
    "b2c_run_device": (_i, [_vp, _u8p, _sz, _sz, _i, _u8p, _sz, _sz, _vp]),
    "b2c_stencil_device": (_i, [_vp, _u8p, _sz, _sz, _i, _vp]),
    "b2c_hysteresis_device": (_i, [_vp, _i, _u8p, _sz, _sz, _vp]),
    "b2c_load_thresh": (_i, [_vp, _vp, _sz]),
    "b2c_run_batch_host": (_i, [_vp, _u8p, _sz, _i, _u8p, _i]),
    "b2c_get_buffer": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_i)]),
    "b2c_download": (_i, [_vp, _i, _vp, _sz]),
    "b2c_copy_view": (_i, [_vp, _vp, _sz, _vp]),
    "b2c_dev_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "b2c_dev_free": (_i, [_vp, _vp]),
    "b2c_dev_upload": (_i, [_vp, _vp, _vp, _sz]),
    "b2c_dev_download": (_i, [_vp, _vp, _vp, _sz]),
    "b2c_sync": (_i, [_vp]),
    "b2c_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "b2c_host_free": (_i, [_vp]),
    "b2c_stream": (_vp, [_vp]),
    "b2c_create_band": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "b2c_band_stencil": (_i, [_vp, _u8p, _sz, _vp]),
    "b2c_band_hysteresis": (_i, [_vp, _vp]),
    "b2c_band_seam_bytes": (_i, [_vp, C.POINTER(_sz)]),
    "b2c_band_seam_record": (_i, [_vp, C.POINTER(_vp)]),
    "b2c_band_seam_solve": (_i, [_vp, _vp, _i, _i, _vp]),
    "b2c_band_status": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "b2c_band_input": (_i, [_vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "b2c_band_p2p_export": (_i, [_vp, _vp]),
    "b2c_band_p2p_open": (_i, [_vp, _vp, _i, _i]),
    "b2c_band_p2p_open_local": (_i, [_vp, C.POINTER(_vp), _i, _i]),
    "b2c_band_p2p_stencil": (_i, [_vp, _vp, _i]),
    "b2c_band_p2p_seam": (_i, [_vp, _vp]),
    "b2c_strerror": (C.c_char_p, [_i]),
    "b2c_last_cuda_error": (C.c_char_p, [_vp]),
    "b2c_version": (C.c_char_p, []),
    "b2c_device_count": (_i, []),
    "b2c_bind_host_to_device": (_i, [_i]),
    "b2c_launch_count": (C.c_longlong, [_vp]),
    "b2c_set_option": (_i, [_vp, C.c_char_p, _i]),
    "b2c_get_info": (_i, [_vp, C.c_char_p]),
    "b2c_synth_frame": (_i, [_i, C.c_uint64, _i, _i, _vp, _sz]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

EXPORTS = tuple(_SIGS)


class B2cError(RuntimeError):
    def __init__(self, status, handle=None, what=""):
        msg = lib.b2c_strerror(status).decode()
        if handle and status == ERR_CUDA:
            msg += " -- " + lib.b2c_last_cuda_error(handle).decode()
        super().__init__(f"{what}: {msg} (status {status})")
        self.status = status


def check(status, handle=None, what="b2c"):
    if status != OK:
        raise B2cError(status, handle, what)
    return status
