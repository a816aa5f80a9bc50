"""Deterministic synthetic BGR8 frames; thin wrapper over b2c_synth_frame (csrc/synth.cpp)."""
import numpy as np

from . import _lib

KINDS = {"scene": 0, "noise": 1, "steps": 2}


def frame(kind, seed, w, h, out=None):
    """Returns an (h, w, 3) uint8 BGR frame."""
    k = KINDS[kind] if isinstance(kind, str) else int(kind)
    if out is None:
        out = np.empty((h, w, 3), np.uint8)
    assert out.dtype == np.uint8 and out.shape == (h, w, 3) and out.strides[2] == 1 and out.strides[1] == 3
    _lib.check(_lib.lib.b2c_synth_frame(k, seed & (2**64 - 1), w, h, out.ctypes.data, out.strides[0]), what="b2c_synth_frame")
    return out


def stream_seed(stream, f):
    """Seed of frame f of stream s."""
    return 0xC0FFEE ^ (stream << 32) ^ f


def batch(kind, n, w, h, stream=0, distinct=None):
    """n frames (n, h, w, 3); only `distinct` different frames are generated, then tiled."""
    distinct = n if distinct is None else min(distinct, n)
    out = np.empty((n, h, w, 3), np.uint8)
    for f in range(distinct):
        frame(kind, stream_seed(stream, f), w, h, out[f])
    for f in range(distinct, n):
        out[f] = out[f % distinct]
    return out


_GIGA_TILE = 4096
_GIGA_PITCH = 3960   # mosaic pitch: NOT a divisor of the band heights (16384 / 2, 4, 8), so that the hard tile-to-tile
                     # edges of the mosaic do not lie exactly on the band seams (a picture-wide edge ON a seam makes
                     # every weak chain along it cross the seam again and again -- the round-based protocol of round 1 needed
                     # 25 global rounds there instead of 3; the seam-graph solve does not care, the picture is kept for comparability)


def giga_rows(y0, y1, w=16384, h=16384):
    """Rows [y0, y1) of the synthetic gigapixel image of BASELINE config 5: a mosaic of four distinct 'scene'
    pictures (tile (ty, tx) shows picture (ty + tx) % 4) at a pitch of 3960 pixels, cropped to w x h.
    Returns (y1-y0, w, 3) uint8."""
    T, Pt = _GIGA_TILE, _GIGA_PITCH
    tiles = {}
    out = np.empty((y1 - y0, w, 3), np.uint8)
    for y in range(y0, y1):
        ty, ry = divmod(y, Pt)
        for tx in range((w + Pt - 1) // Pt):
            k = (ty + tx) % 4
            if k not in tiles:
                tiles[k] = frame("scene", stream_seed(1000, k), T, T)
            x0 = tx * Pt
            n = min(Pt, w - x0)
            out[y - y0, x0:x0 + n] = tiles[k][ry, :n]
    return out
