"""cudacam_b200 -- B200-native Canny edge detector behind the class surface of axoloto/CudaCam's src/cvp.

The product is cudacam_b200/libb200canny.so (hand-written CUDA for sm_100a + a C ABI, include/b200canny.h);
this package is the Python view of that ABI used by the tests and the bench.  No CPU fallback exists.
"""
from ._lib import B2cError, LIB_PATH  # noqa: F401
from .canny import CANNY_STAGES, CannyEdge, CannyStage, CvPipeline  # noqa: F401
from . import synth  # noqa: F401
