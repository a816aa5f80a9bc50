"""Host-side mirror of the reference's Canny classes on top of the C ABI.

`CannyEdge` follows cvp::cuda::CannyEdge (reference src/cvp/cannyEdgeH.hpp:17-110): construct for given frame
dimensions, `run(frame, final_stage)` per frame, low/high threshold setters with the same clamping
(cannyEdgeH.hpp:25-29), kernel-profiling toggle (:31-32) -- plus the explicit accessors for the intermediate
buffers that the reference only exposes through its `finalStage` switch (cannyEdgeH.cu:169-207).
`CvPipeline` follows cvp::cvPipeline (src/cvp/cvPipeline.{hpp,cpp}): same argument checks, returns bool.
The C++ twin of this file is include/b200canny.hpp; both are thin: all work happens behind the C ABI in CUDA.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import lib, check


class CannyStage(enum.IntEnum):
    """cvp::CannyStage (src/cvp/define.hpp:9-17)."""
    MONO = 0
    GAUSSIAN = 1
    GRADIENT = 2
    NMS = 3
    THRESH = 4
    HYSTER = 5


#: src/cvp/define.hpp:27-34
CANNY_STAGES = {
    CannyStage.MONO: "1/6 Mono Conversion",
    CannyStage.GAUSSIAN: "2/6 Gaussian Noise Removal",
    CannyStage.GRADIENT: "3/6 Gradient Computation",
    CannyStage.NMS: "4/6 Non Maximum Suppression",
    CannyStage.THRESH: "5/6 Double Threshold",
    CannyStage.HYSTER: "6/6 Hysteresis",
}


PLANAR_BGR8 = 0x103   # include/b200canny.h B2C_PLANAR_BGR8


class CannyEdge:
    def __init__(self, width, height, channels=3, device=0, max_batch=1, planar=False):
        """planar=True (with channels=3): frames are three planes B, G, R -- arrays of shape (3, h, w)."""
        self.width, self.height, self.channels, self.device, self.max_batch = width, height, channels, device, max_batch
        self.planar = bool(planar)
        if self.planar and channels != 3:
            raise ValueError("planar input is BGR8")
        h = C.c_void_p()
        check(lib.b2c_create(C.byref(h), device, width, height, PLANAR_BGR8 if self.planar else channels, max_batch), what="b2c_create")
        self._h = h

    # -- lifetime ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib.b2c_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- thresholds / profiling (cannyEdgeH.hpp:25-32) --------------------------------------------------------
    def setLowThreshold(self, low):
        check(lib.b2c_set_low_threshold(self._h, int(low) & 0xFF), self._h)

    def getLowThreshold(self):
        return lib.b2c_get_low_threshold(self._h)

    def setHighThreshold(self, high):
        check(lib.b2c_set_high_threshold(self._h, int(high) & 0xFF), self._h)

    def getHighThreshold(self):
        return lib.b2c_get_high_threshold(self._h)

    def enableKernelProfiling(self, on):
        check(lib.b2c_enable_profiling(self._h, 1 if on else 0), self._h)

    def isKernelProfilingEnabled(self):
        return bool(lib.b2c_is_profiling_enabled(self._h))

    def lastTimings(self):
        """dict of ms: upload, stencil, hysteresis, output, total; rounds = on-device hysteresis passes (always 1)."""
        v = (C.c_float * 6)()
        check(lib.b2c_last_timings(self._h, v, 6), self._h, "b2c_last_timings")
        return dict(upload=v[0], stencil=v[1], hysteresis=v[2], output=v[3], total=v[4], rounds=int(v[5]))

    # -- run (cannyEdgeH.cu:49-120) -----------------------------------------------------------------------------
    def run(self, frame, final_stage=CannyStage.HYSTER):
        """frame: (h, w, channels) uint8 host array -- BGR8, BGRA8 or GRAY8 ((h, w) also accepted) as given to the
        constructor; rows may be strided, like cv::Mat::step."""
        f = np.asarray(frame)
        if self.planar:   # (3, h, w): planes B, G, R, rows may be strided, planes must follow each other
            if f.dtype != np.uint8 or f.shape != (3, self.height, self.width) or f.strides[2] != 1 or f.strides[0] != f.strides[1] * self.height:
                raise ValueError("planar frame must be a (3, h, w) uint8 array of consecutive planes")
            check(lib.b2c_run(self._h, f.ctypes.data, f.strides[1], int(final_stage)), self._h, "b2c_run")
            return
        if f.ndim == 2:
            f = f[:, :, None]
        ch = self.channels
        if f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != ch or f.strides[2] != 1 or f.strides[1] != ch:
            raise ValueError(f"frame must be an (h, w, {ch}) uint8 array with packed pixels")
        if f.shape[0] != self.height or f.shape[1] != self.width:
            raise _lib.B2cError(_lib.ERR_SIZE, what="run")
        check(lib.b2c_run(self._h, f.ctypes.data, f.strides[0], int(final_stage)), self._h, "b2c_run")

    def copy_view(self, dev_dst, pitch=0, stream=None):
        """The GL-free half of _sendOutputToOpenGL (cannyEdgeH.cu:154-212): the u8 picture of the stage the last run stopped
        at, device to device into dev_dst (an int address, e.g. a mapped PBO), rows `pitch` bytes apart (0 = width)."""
        check(lib.b2c_copy_view(self._h, dev_dst, pitch, stream), self._h, "b2c_copy_view")

    def run_device(self, dev_ptr, row_stride, frame_stride, n, edges_ptr=None, edges_pitch=0, edges_frame_stride=0, stream=None):
        check(lib.b2c_run_device(self._h, dev_ptr, row_stride, frame_stride, n, edges_ptr, edges_pitch, edges_frame_stride, stream), self._h, "b2c_run_device")

    def run_batch(self, frames, packed_bits=False, out=None):
        """frames: (n, h, w, channels) uint8 host array -> (n, h, w) uint8 edge maps (or (n, h, ceil(w/32)) uint32 bit maps)."""
        f = np.asarray(frames)
        shape = (3, self.height, self.width) if self.planar else (self.height, self.width, self.channels)
        if f.dtype != np.uint8 or f.ndim != 4 or f.shape[1:] != shape or not f.flags.c_contiguous:
            raise ValueError("frames must be a C-contiguous (n, h, w, channels) -- planar: (n, 3, h, w) -- uint8 array")
        n = f.shape[0]
        if out is None:
            out = np.empty((n, self.height, (self.width + 31) // 32), np.uint32) if packed_bits else np.empty((n, self.height, self.width), np.uint8)
        check(lib.b2c_run_batch_host(self._h, f.ctypes.data, self.width * (1 if self.planar else self.channels), n, out.ctypes.data, 1 if packed_bits else 0), self._h, "b2c_run_batch_host")
        return out

    # -- accessors ----------------------------------------------------------------------------------------------
    def _download(self, buf, dtype, cols=None):
        out = np.empty((self.height, self.width if cols is None else cols), dtype)
        check(lib.b2c_download(self._h, buf, out.ctypes.data, out.strides[0]), self._h, "b2c_download")
        return out

    def mono(self):
        return self._download(_lib.BUF_MONO, np.uint8)

    def blur(self):
        return self._download(_lib.BUF_BLUR, np.uint8)

    def gradient(self):
        return self._download(_lib.BUF_GRAD, np.float32)

    def nms(self):
        return self._download(_lib.BUF_NMS, np.uint8)

    def thresh(self):
        return self._download(_lib.BUF_THRESH, np.uint8)

    def edges(self):
        return self._download(_lib.BUF_EDGES, np.uint8)

    def view(self):
        """What the reference copies into its GL PBO for the stage of the last run (cannyEdgeH.cu:154-212)."""
        return self._download(_lib.BUF_VIEW, np.uint8)

    def map2(self):
        return self._download(_lib.BUF_MAP2, np.uint32, (self.width + 15) // 16)

    def bits(self):
        return self._download(_lib.BUF_BITS, np.uint32, (self.width + 31) // 32)

    def sync(self):
        check(lib.b2c_sync(self._h), self._h)

    def set_option(self, name, value):
        check(lib.b2c_set_option(self._h, name.encode(), int(value)), self._h, "b2c_set_option")

    def info(self, name):
        return lib.b2c_get_info(self._h, name.encode())

    @property
    def launches(self):
        return lib.b2c_launch_count(self._h)


class CvPipeline:
    """cvp::cvPipeline (src/cvp/cvPipeline.hpp:20-39).  `pbo` is accepted and ignored (no GL here): the bytes the
    reference would put in the PBO are returned by `output()`."""

    def __init__(self, pbo, inputImageCols, inputImageRows, inputImageNbChannels, device=0):
        self._edge = CannyEdge(inputImageCols, inputImageRows, inputImageNbChannels, device)
        self._pbo = pbo

    def process(self, inputImage, finalStage):
        # src/cvp/cvPipeline.cpp:19-41: null impl / empty frame / wrong type -> false
        if self._edge is None or inputImage is None:
            return False
        f = np.asarray(inputImage)
        if f.ndim == 2:
            f = f[:, :, None]
        if f.size == 0 or f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != self._edge.channels:
            return False
        self._edge.run(f, finalStage)
        return True

    def output(self):
        return self._edge.view()

    def setLowThreshold(self, low):
        self._edge.setLowThreshold(low)

    def getLowThreshold(self):
        return self._edge.getLowThreshold()

    def setHighThreshold(self, high):
        self._edge.setHighThreshold(high)

    def getHighThreshold(self):
        return self._edge.getHighThreshold()

    def enableCudaProfiling(self, on):
        self._edge.enableKernelProfiling(on)

    def isCudaProfilingEnabled(self):
        return self._edge.isKernelProfilingEnabled()
