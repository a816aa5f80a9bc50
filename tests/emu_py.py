"""TEST-ONLY: loads tests/emu/libemu.so, the product's kernel sources compiled for the CPU through
tests/emu/cuda_emu.h.  Lets `-m "not gpu"` tests check kernel logic against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_SO = os.path.join(EMU_DIR, "libemu.so")
# B2C_EMU_SO=/path/to/libemu_asan.so (built by tools/emu_asan.sh with -fsanitize=address, run under LD_PRELOAD=libasan)
# turns the emulator tests into an out-of-bounds check of the kernels' indexing: compute-sanitizer is closed on the GPU pool
EMU_OVERRIDE = os.environ.get("B2C_EMU_SO")
_vp, _i, _ll = C.c_void_p, C.c_int, C.c_longlong


def build(force=False):
    srcs = [os.path.join(EMU_DIR, "emu_main.cpp"), os.path.join(EMU_DIR, "cuda_emu.h")]
    csrc = os.path.join(os.path.dirname(HERE), "cudacam_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not force and os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", "-fPIC", "-shared", "-fvisibility=hidden", "-Wno-unused",
                           "-DB2C_EMU_FUSED", "-I", EMU_DIR, "-o", EMU_SO, srcs[0]])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not EMU_OVERRIDE:
            build()
        _lib = C.CDLL(EMU_OVERRIDE or EMU_SO)
        _lib.emu_stencil.restype = _i
        _lib.emu_stencil.argtypes = [_i, _vp, _ll, _ll, _i, _i, _i, _i, _i, C.c_uint, C.c_uint, _vp, _vp, _vp, _vp, _vp, _vp]
        _lib.emu_set_channels.restype = None
        _lib.emu_set_channels.argtypes = [_i]
        _lib.emu_set_spread.restype = None
        _lib.emu_set_spread.argtypes = [_i]
        _lib.emu_set_plane_stride.restype = None
        _lib.emu_set_plane_stride.argtypes = [_ll]
        _lib.emu_hysteresis.restype = _i
        _lib.emu_hysteresis.argtypes = [_vp, _i, _i, _i, _vp, _vp]
        _lib.emu_band_create.restype = _vp
        _lib.emu_band_create.argtypes = [_i, _i, _i, _i]
        _lib.emu_band_destroy.restype = None
        _lib.emu_band_destroy.argtypes = [_vp]
        _lib.emu_seam_words.restype = _i
        _lib.emu_seam_words.argtypes = [_i]
        _lib.emu_band_hysteresis.restype = None
        _lib.emu_band_hysteresis.argtypes = [_vp, _vp, _vp]
        _lib.emu_band_unresolved_words.restype = _i
        _lib.emu_band_unresolved_words.argtypes = [_vp]
        _lib.emu_band_solve.restype = _i
        _lib.emu_band_solve.argtypes = [_vp, _vp, _i, _i]
        _lib.emu_band_edges.restype = None
        _lib.emu_band_edges.argtypes = [_vp, _vp]
    return _lib


def stencil(bgr, lo=10, hi=40, impl=1, stages=False, y0=0, h_glob=None, rows=None, row0=0):
    """bgr (h, w, 3) or (n, h, w, 3).  In band mode pass the whole image plus row0/rows/y0/h_glob."""
    a = np.ascontiguousarray(bgr, np.uint8)
    if a.ndim == 3:
        a = a[None]
    n, hh, w, ch = a.shape
    h = hh if rows is None else rows
    h_glob = hh if h_glob is None else h_glob
    gpr = (w + 15) // 16
    map2 = np.zeros((n, h, gpr), np.uint32)
    out = dict(map2=map2)
    ptrs = [None] * 5
    if stages:
        out.update(mono=np.zeros((h, w), np.uint8), blur=np.zeros((h, w), np.uint8), grad=np.zeros((h, w), np.float32),
                   nms=np.zeros((h, w), np.uint8), thresh=np.zeros((h, w), np.uint8))
        ptrs = [out[k].ctypes.data for k in ("mono", "blur", "grad", "nms", "thresh")]
    base = a.ctypes.data + row0 * a.strides[1]
    lib().emu_set_channels(ch)
    try:
        rc = lib().emu_stencil(impl, base, a.strides[1], a.strides[0], w, h, y0, h_glob, n, lo, hi, map2.ctypes.data, *ptrs)
    finally:
        lib().emu_set_channels(3)
    assert rc == 0, rc
    return out


class Band:
    """One row band with retained planes and union-find forest (the emulated twin of a b2c band handle)."""

    def __init__(self, w, h, ucap=0, force_global=False):
        self.w, self.h = w, h
        self._h = lib().emu_band_create(w, h, ucap, 1 if force_global else 0)
        self.seam_words = lib().emu_seam_words(w)
        self.record = np.zeros(self.seam_words, np.uint32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_band_destroy(self._h)
            self._h = None

    def hysteresis(self, map2):
        """Band-local pass; leaves the seam record in self.record."""
        m = np.ascontiguousarray(map2, np.uint32)
        lib().emu_band_hysteresis(self._h, m.ctypes.data, self.record.ctypes.data)

    def unresolved_words(self):
        return lib().emu_band_unresolved_words(self._h)

    def solve(self, all_records, world, rank):
        a = np.ascontiguousarray(all_records, np.uint32)
        return lib().emu_band_solve(self._h, a.ctypes.data, world, rank)

    def edges(self):
        out = np.zeros((self.h, self.w), np.uint8)
        lib().emu_band_edges(self._h, out.ctypes.data)
        return out


def hysteresis(map2, w, want_edges=True, spread=8):
    """(n, h, gpr) or (h, gpr) 2-bit maps -> (u8 edge maps or None, edge bit planes).  spread: warps per tile the work
    items of k_uf_tile are dealt to (the product picks 8 for small batches, 4 for big ones)."""
    m = np.ascontiguousarray(map2, np.uint32)
    if m.ndim == 2:
        m = m[None]
    n, h, _ = m.shape
    edges = np.zeros((n, h, w), np.uint8) if want_edges else None
    bits = np.zeros((n, h, (w + 31) // 32), np.uint32)
    lib().emu_set_spread(spread)
    try:
        rc = lib().emu_hysteresis(m.ctypes.data, w, h, n, None if edges is None else edges.ctypes.data, bits.ctypes.data)
    finally:
        lib().emu_set_spread(8)
    assert rc == 0
    return edges, bits


def stencil_raw(buf, row0, w, h, lo=10, hi=40, impl=0, y0=0, h_glob=None, channels=3, plane_stride=0):
    """buf: 2-D uint8 array of padded rows (w * channels bytes used per row); the frame starts at row `row0`.  Returns
    the 2-bit map or None if the implementation is not available in the emulator build.  plane_stride != 0: planar BGR8
    (the G and R planes lie plane_stride and 2 * plane_stride bytes behind the B plane)."""
    h_glob = h if h_glob is None else h_glob
    map2 = np.zeros((h, (w + 15) // 16), np.uint32)
    lib().emu_set_channels(channels)
    lib().emu_set_plane_stride(plane_stride)
    try:
        rc = lib().emu_stencil(impl, buf.ctypes.data + row0 * buf.strides[0], buf.strides[0], 0, w, h, y0, h_glob, 1, lo, hi, map2.ctypes.data, None, None, None, None, None)
    finally:
        lib().emu_set_channels(3)
        lib().emu_set_plane_stride(0)
    return map2 if rc == 0 else None
