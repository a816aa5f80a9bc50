"""ctypes view of the checkers under oracle/ (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcvpref.so")

_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int


def _load_oracle():
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_canny.restype = _i
    lib.oracle_canny.argtypes = [_vp, _sz, _i, _i, C.c_uint8, C.c_uint8] + [_vp] * 7
    lib.oracle_canny_ch.restype = _i
    lib.oracle_canny_ch.argtypes = [_vp, _sz, _i, _i, _i, C.c_uint8, C.c_uint8] + [_vp] * 7
    lib.oracle_hysteresis.restype = None
    lib.oracle_hysteresis.argtypes = [_vp, _i, _i, _vp]
    lib.oracle_hysteresis_launches.restype = _i
    lib.oracle_hysteresis_launches.argtypes = [_vp, _i, _i, _i, _vp, C.POINTER(_i)]
    lib.oracle_sector.restype = _i
    lib.oracle_sector.argtypes = [_i, _i]
    lib.oracle_gauss_kernel.restype = None
    lib.oracle_gauss_kernel.argtypes = [_vp]
    lib.oracle_float2uchar.restype = None
    lib.oracle_float2uchar.argtypes = [_vp, _sz, _vp]
    lib.oracle_set_low.restype = None
    lib.oracle_set_low.argtypes = [_vp, _vp, C.c_uint8]
    lib.oracle_set_high.restype = None
    lib.oracle_set_high.argtypes = [_vp, _vp, C.c_uint8]
    return lib


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        _oracle = _load_oracle()
    return _oracle


def canny(bgr, low=10, high=40, want_edges=True):
    """Runs the CPU restatement.  Returns dict(mono, blur, grad, sector, nms, thresh, edges)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    if bgr.ndim == 2:
        bgr = bgr[:, :, None]
    h, w, ch = bgr.shape
    out = dict(mono=np.empty((h, w), np.uint8), blur=np.empty((h, w), np.uint8), grad=np.empty((h, w), np.float32),
               sector=np.empty((h, w), np.uint8), nms=np.empty((h, w), np.uint8), thresh=np.empty((h, w), np.uint8),
               edges=np.empty((h, w), np.uint8))
    rc = oracle().oracle_canny_ch(bgr.ctypes.data, bgr.strides[0], w, h, ch, low, high, out["mono"].ctypes.data, out["blur"].ctypes.data,
                               out["grad"].ctypes.data, out["sector"].ctypes.data, out["nms"].ctypes.data, out["thresh"].ctypes.data,
                               out["edges"].ctypes.data if want_edges else None)
    assert rc == 0
    return out


def hysteresis(thresh):
    t = np.ascontiguousarray(thresh, np.uint8)
    out = np.empty_like(t)
    oracle().oracle_hysteresis(t.ctypes.data, t.shape[1], t.shape[0], out.ctypes.data)
    return out


def hysteresis_launches(thresh, max_iters=100):
    t = np.ascontiguousarray(thresh, np.uint8)
    flag = C.c_int(0)
    state = np.empty_like(t)
    it = oracle().oracle_hysteresis_launches(t.ctypes.data, t.shape[1], t.shape[0], max_iters, state.ctypes.data, C.byref(flag))
    return it, flag.value, state


def float2uchar(grad):
    g = np.ascontiguousarray(grad, np.float32)
    out = np.empty(g.shape, np.uint8)
    oracle().oracle_float2uchar(g.ctypes.data, g.size, out.ctypes.data)
    return out


def thresh_to_map2(thresh):
    """u8 {0,128,255} -> the packed 2-bit map layout of the product (u32 per 16 px: bits 0-15 strong, 16-31 weak)."""
    h, w = thresh.shape
    g = (w + 15) // 16
    pad = np.zeros((h, g * 16), np.uint8)
    pad[:, :w] = thresh
    s = (pad == 255).reshape(h, g, 16).astype(np.uint32)
    k = (pad == 128).reshape(h, g, 16).astype(np.uint32)
    sh = np.arange(16, dtype=np.uint32)
    return ((s << sh).sum(2) | ((k << sh).sum(2) << 16)).astype(np.uint32)


def edges_to_bits(edges):
    h, w = edges.shape
    g = (w + 31) // 32
    pad = np.zeros((h, g * 32), np.uint8)
    pad[:, :w] = edges
    b = (pad == 255).reshape(h, g, 32).astype(np.uint64)
    return (b << np.arange(32, dtype=np.uint64)).sum(2).astype(np.uint32)


# ---- the reference's own kernels (GPU only) -------------------------------------------------------------------
class CvpRef:
    """Driver of the UNMODIFIED reference kernels (oracle/ref_harness.cu -> oracle/_ref/libcvpref.so)."""
    NAMES = ["mono", "blur", "sobelX", "sobelY", "grad", "slope", "nms", "thresh", "hyster", "pbo"]

    def __init__(self, w, h):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.cvpref_create.restype = _i
        L.cvpref_create.argtypes = [C.POINTER(_vp), _i, _i]
        L.cvpref_destroy.argtypes = [_vp]
        L.cvpref_set_thresholds.argtypes = [_vp, C.c_uint8, C.c_uint8]
        L.cvpref_enable_profiling.argtypes = [_vp, _i]
        L.cvpref_run.restype = _i
        L.cvpref_run.argtypes = [_vp, _vp, _sz, _i]
        L.cvpref_download.restype = _i
        L.cvpref_download.argtypes = [_vp, _i, _vp]
        L.cvpref_info.restype = _i
        L.cvpref_info.argtypes = [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_float)]
        L.cvpref_gradslope_table.restype = _i
        L.cvpref_gradslope_table.argtypes = [_vp, _vp]
        L.cvpref_nms_raw.restype = _i
        L.cvpref_nms_raw.argtypes = [_vp, _vp, _i, _i, _vp]
        self.w, self.h = w, h
        self._h = _vp()
        rc = L.cvpref_create(C.byref(self._h), w, h)
        if rc != 0:
            raise RuntimeError(f"cvpref_create failed: {rc}")

    def close(self):
        if self._h:
            self.lib.cvpref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_thresholds(self, lo, hi):
        self.lib.cvpref_set_thresholds(self._h, lo, hi)

    def enable_profiling(self, on):
        self.lib.cvpref_enable_profiling(self._h, 1 if on else 0)

    def run(self, bgr, final_stage=5):
        assert bgr.dtype == np.uint8 and bgr.shape[:2] == (self.h, self.w)
        rc = self.lib.cvpref_run(self._h, bgr.ctypes.data, bgr.strides[0], final_stage)
        if rc != 0:
            raise RuntimeError(f"cvpref_run failed: {rc}")

    def get(self, name):
        i = self.NAMES.index(name)
        dt = np.float32 if name in ("sobelX", "sobelY", "grad", "slope") else np.uint8
        out = np.empty((self.h, self.w), dt)
        rc = self.lib.cvpref_download(self._h, i, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cvpref_download failed: {rc}")
        return out

    def info(self):
        it, fl = C.c_int(0), C.c_int(0)
        ms = (C.c_float * 6)()
        self.lib.cvpref_info(self._h, C.byref(it), C.byref(fl), ms)
        return it.value, fl.value, list(ms)

    def gradslope_table(self):
        n = 2041
        g = np.empty((n, n), np.float32)
        s = np.empty((n, n), np.float32)
        rc = self.lib.cvpref_gradslope_table(g.ctypes.data, s.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cvpref_gradslope_table failed: {rc}")
        return g, s

    def nms_raw(self, grad, slope):
        h, w = grad.shape
        out = np.empty((h, w), np.uint8)
        rc = self.lib.cvpref_nms_raw(np.ascontiguousarray(grad, np.float32).ctypes.data, np.ascontiguousarray(slope, np.float32).ctypes.data, w, h, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cvpref_nms_raw failed: {rc}")
        return out
