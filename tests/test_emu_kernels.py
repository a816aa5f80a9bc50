"""Kernel LOGIC on the CPU: the product's .cuh kernels compiled for the host through tests/emu/cuda_emu.h
(one OS thread per CUDA thread, real barriers) and compared with the oracle.  This is not the product path
(that is CUDA only, covered by the -m gpu tests); it catches indexing / synchronisation bugs without a GPU."""
import numpy as np
import pytest

import emu_py as E
import oracle_py as O
from cudacam_b200 import synth

CASES = [("scene", 200, 150, 7, 10, 40), ("noise", 97, 61, 8, 10, 40), ("steps", 130, 70, 9, 17, 43), ("scene", 16, 5, 3, 10, 40), ("noise", 1, 1, 1, 0, 0)]


@pytest.mark.parametrize("kind,w,h,seed,lo,hi", CASES)
def test_tile_stencil_all_stages(kind, w, h, seed, lo, hi):
    f = synth.frame(kind, seed, w, h)
    r = O.canny(f, lo, hi)
    e = E.stencil(f, lo, hi, impl=1, stages=True)
    for k in ("mono", "blur", "nms", "thresh"):
        assert np.array_equal(e[k], r[k]), k
    assert np.array_equal(e["grad"].view(np.uint32), r["grad"].view(np.uint32))
    assert np.array_equal(e["map2"][0], O.thresh_to_map2(r["thresh"]))


MARCH_CASES = [("scene", 200, 150, 7, 10, 40), ("noise", 96, 61, 8, 10, 40), ("steps", 136, 70, 9, 17, 43), ("scene", 16, 5, 3, 10, 40),
               ("steps", 480, 130, 2, 3, 200), ("scene", 8, 8, 1, 0, 0), ("noise", 488, 64, 3, 10, 40), ("scene", 720, 64, 3, 10, 40),
               # widths that are not a multiple of 8: the last lane of the strip is only partly inside the image
               ("scene", 201, 50, 7, 10, 40), ("noise", 97, 61, 8, 10, 40), ("steps", 243, 40, 9, 17, 43), ("scene", 13, 9, 3, 10, 40), ("noise", 487, 30, 3, 10, 40)]


def _padded(f, ch=3, top=4):
    """Rows padded the way the host driver pads them: whole 8-pixel lanes backed by memory, 16-byte aligned."""
    h, w = f.shape[:2]
    stride = ((w + 7) // 8 * 8 * ch + 15) // 16 * 16
    buf = np.zeros((h + 2 * top, stride), np.uint8)
    buf[top:top + h, :w * ch] = f.reshape(h, w * ch)
    return buf


@pytest.mark.parametrize("rb", [8, 20, 44], ids=["rb8", "rb20", "rb44"])
@pytest.mark.parametrize("kind,w,h,seed,lo,hi", MARCH_CASES)
def test_march_stencil_map(kind, w, h, seed, lo, hi, rb):
    """impl 100 + rb = marching two-warp pipeline kernel with rb rows per band."""
    f = synth.frame(kind, seed, w, h)
    r = O.canny(f, lo, hi, want_edges=False)
    e = E.stencil_raw(_padded(f), 4, w, h, lo, hi, impl=100 + rb)
    assert e is not None and np.array_equal(e, O.thresh_to_map2(r["thresh"]))


@pytest.mark.parametrize("kind,w,h,seed", [("scene", 200, 150, 7), ("noise", 97, 61, 8), ("steps", 130, 70, 9), ("scene", 1100, 40, 5)])
def test_hysteresis_kernel(kind, w, h, seed):
    f = synth.frame(kind, seed, w, h)
    r = O.canny(f)
    edges, bits = E.hysteresis(O.thresh_to_map2(r["thresh"]), w)
    assert np.array_equal(edges[0], r["edges"])
    assert np.array_equal(bits[0], O.edges_to_bits(r["edges"]))
    # bit plane only (the u8 map is optional in the product: b2c_run_batch_* with edges == null)
    none, bits2 = E.hysteresis(O.thresh_to_map2(r["thresh"]), w, want_edges=False)
    assert none is None and np.array_equal(bits2, bits)


@pytest.mark.parametrize("spread", [1, 2, 4])
def test_hysteresis_item_order_variants(spread):
    """k_uf_tile deals its work items to 1, 2, 4 or 8 warps (8 is what every other test runs): same result."""
    rng = np.random.default_rng(11)
    t = np.where(rng.random((70, 530)) < 0.35, 128, 0).astype(np.uint8)
    t[rng.random(t.shape) < 0.002] = 255
    edges, _ = E.hysteresis(O.thresh_to_map2(t), 530, spread=spread)
    assert np.array_equal(edges[0], O.hysteresis(t))


@pytest.mark.parametrize("dens", [0.1, 0.3, 0.5])
def test_hysteresis_unionfind_random_maps(dens):
    rng = np.random.default_rng(int(dens * 10))
    t = np.where(rng.random((90, 131)) < dens, 128, 0).astype(np.uint8)
    t[rng.random(t.shape) < 0.002] = 255
    edges, _ = E.hysteresis(O.thresh_to_map2(t), 131)
    assert np.array_equal(edges[0], O.hysteresis(t))


def test_hysteresis_long_chain_and_batch():
    # a weak spiral seeded by a single strong pixel: worst case for tile-local propagation
    w, h = 70, 40
    t = np.zeros((h, w), np.uint8)
    t[2, 2:w - 2] = 128
    t[2:h - 2, w - 3] = 128
    t[h - 3, 4:w - 2] = 128
    t[6:h - 2, 4] = 128
    t[6, 4:w - 6] = 128
    t[2, 2] = 255
    t[20, 30] = 128   # isolated weak pixel: must vanish
    m = np.stack([O.thresh_to_map2(t), O.thresh_to_map2(np.zeros_like(t))])
    edges, bits = E.hysteresis(m, w)
    assert np.array_equal(edges[0], O.hysteresis(t)) and edges[0][20, 30] == 0 and edges[0][6, w - 7] == 255
    assert not edges[1].any()


def test_band_mode_stencil_equals_whole_image():
    w, h = 150, 90
    f = synth.frame("scene", 21, w, h)
    whole = E.stencil(f, impl=1)["map2"][0]
    y0, rows = 30, 25
    band = E.stencil(f, impl=1, y0=y0, h_glob=h, rows=rows, row0=y0)["map2"][0]
    assert np.array_equal(band, whole[y0:y0 + rows])


@pytest.mark.parametrize("impl", [132], ids=["march"])
@pytest.mark.parametrize("v", [0, 3, 100, 255])
def test_march_flat_picture_takes_dense_replay(v, impl):
    """Flat regions make S % 159 == 0 for every pixel: the work list overflows and the dense replay runs."""
    f = np.full((70, 248, 3), v, np.uint8)
    f[30:40, 100:140] = (v + 60) % 256
    r = O.canny(f, 10, 40, want_edges=False)
    buf = np.zeros((78, 752), np.uint8)
    buf[4:74, :744] = f.reshape(70, -1)
    e = E.stencil_raw(buf, 4, 248, 70, 10, 40, impl=impl)
    assert e is not None and np.array_equal(e, O.thresh_to_map2(r["thresh"]))


@pytest.mark.parametrize("impl", [120, 132], ids=["march20", "march32"])
def test_march_band_mode_equals_whole_image(impl):
    w, h = 248, 150
    f = synth.frame("scene", 23, w, h)
    whole = O.thresh_to_map2(O.canny(f, want_edges=False)["thresh"])
    buf = np.zeros((h, 752), np.uint8)
    buf[:, :w * 3] = f.reshape(h, -1)
    for y0, rows in ((0, 50), (50, 61), (111, 39)):
        band = E.stencil_raw(buf, y0, w, rows, impl=impl, y0=y0, h_glob=h)
        assert np.array_equal(band, whole[y0:y0 + rows]), (y0, rows)


def test_hysteresis_wide_dense_map():
    """8 tile columns of dense clutter: long border lists between the 32-row x 256-pixel tiles."""
    rng = np.random.default_rng(5)
    t = np.where(rng.random((80, 2048)) < 0.2, 128, 0).astype(np.uint8)
    t[rng.random(t.shape) < 0.001] = 255
    edges, _ = E.hysteresis(O.thresh_to_map2(t), 2048)
    assert np.array_equal(edges[0], O.hysteresis(t))


@pytest.mark.parametrize("dens,seed", [(0.08, 1), (0.25, 2), (0.6, 3)])
def test_hysteresis_unionfind4_many_tiles(dens, seed):
    """Random weak maps spanning several 32-row x 256-pixel tiles in both directions, sparse strong seeds: exercises the
    tile-local forests, the border lists (top rows, left / right word columns, NW / NE across tile corners) and resolve."""
    rng = np.random.default_rng(seed)
    w, h = 600, 75
    t = np.where(rng.random((h, w)) < dens, 128, 0).astype(np.uint8)
    t[rng.random(t.shape) < 0.0015] = 255
    # long chains along and across the tile borders
    t[31:33, 100:500] = np.where(t[31:33, 100:500] == 255, 255, 128)
    t[5:70, 255:257] = np.where(t[5:70, 255:257] == 255, 255, 128)
    edges, bits = E.hysteresis(O.thresh_to_map2(t), w)
    want = O.hysteresis(t)
    assert np.array_equal(edges[0], want)
    assert np.array_equal(bits[0], O.edges_to_bits(want))


def test_hysteresis_unionfind4_diagonal_staircase_across_tiles():
    """A one-pixel diagonal staircase crosses tile corners (NW / NE unions across both a row and a word boundary)."""
    w, h = 520, 70
    t = np.zeros((h, w), np.uint8)
    for i in range(60):
        t[5 + i, 225 + i] = 128          # down-right through (32, 256)
        t[5 + i, 290 - i] = 128          # down-left through (32 + ..., 256)
    t[5, 225] = 255
    t[64, 231] = 255
    edges, _ = E.hysteresis(O.thresh_to_map2(t), w)
    assert np.array_equal(edges[0], O.hysteresis(t))


@pytest.mark.parametrize("w,h,rb", [(8, 3, 16), (240, 36, 36), (248, 37, 26), (480, 12, 6), (1000, 23, 16), (241, 25, 8), (239, 13, 20), (249, 50, 44), (9, 1, 8)])
def test_march_geometry_corner_cases(w, h, rb):
    """Strip / band boundaries of the marching kernel: widths around multiples of 240 and 8, heights around the block size."""
    f = synth.frame("scene", 40 + w, w, h)
    r = O.canny(f, 10, 40, want_edges=False)
    e = E.stencil_raw(_padded(f), 4, w, h, 10, 40, impl=100 + rb)
    assert e is not None and np.array_equal(e, O.thresh_to_map2(r["thresh"]))


def test_march_long_band_many_blocks():
    """One band of many 12-row blocks: the producer / consumer hand-over of the blur ring is exercised dozens of times,
    on noise (every lane busy in the sparse stages) so that the two warps drift against each other."""
    w, h = 248, 300
    f = synth.frame("noise", 77, w, h)
    r = O.canny(f, 10, 40, want_edges=False)
    e = E.stencil_raw(_padded(f), 4, w, h, 10, 40, impl=100 + 296)
    assert e is not None and np.array_equal(e, O.thresh_to_map2(r["thresh"]))


def _to_format(f, ch):
    """BGR frame -> GRAY8 (green channel as the gray picture) or BGRA8 (random alpha, must be ignored)."""
    if ch == 1:
        return np.ascontiguousarray(f[:, :, 1:2])
    a = np.random.default_rng(1).integers(0, 256, f.shape[:2] + (1,), dtype=np.uint8)
    return np.ascontiguousarray(np.concatenate([f, a], axis=2))


@pytest.mark.parametrize("ch", [1, 4])
@pytest.mark.parametrize("kind,w,h,seed", [("scene", 200, 60, 7), ("steps", 256, 45, 9), ("noise", 96, 30, 8)])
def test_other_input_formats(kind, w, h, seed, ch):
    """GRAY8 and BGRA8 front ends of the marching and the tile kernel; BGRA must equal BGR, gray must
    equal the oracle run on the gray picture."""
    f = synth.frame(kind, seed, w, h)
    g = _to_format(f, ch)
    r = O.canny(g)
    if ch == 4:
        assert np.array_equal(r["thresh"], O.canny(f)["thresh"])
    e = E.stencil(g, impl=1, stages=True)
    for k in ("mono", "blur", "nms", "thresh"):
        assert np.array_equal(e[k], r[k]), k
    m = E.stencil_raw(_padded(g, ch), 4, w, h, impl=126, channels=ch)
    assert m is not None and np.array_equal(m, O.thresh_to_map2(r["thresh"]))


def test_planar_bgr_tile_kernel():
    """Planar BGR8: the tile kernel reads the three planes; the marching kernel declines (returns None -> tile path)."""
    w, h = 150, 70
    f = synth.frame("scene", 12, w, h)
    r = O.canny(f)
    pitch = 160
    buf = np.zeros((3 * h, pitch), np.uint8)
    buf[:, :w] = f.transpose(2, 0, 1).reshape(3 * h, w)
    m = E.stencil_raw(buf, 0, w, h, impl=1, plane_stride=pitch * h)
    assert m is not None and np.array_equal(m, O.thresh_to_map2(r["thresh"]))
    assert E.stencil_raw(buf, 0, w, h, impl=120, plane_stride=pitch * h) is None


def test_hysteresis_tile_corner_cases():
    """Structured cases around the corner of the 32-row x 256-pixel tiles of k_uf_tile (y = 32, x = 256):
    contacts with promoted pixels of a neighbour tile straight and diagonally across the corner, a chain through four
    tiles seeded in the last one, closed blobs, open components that die, the image's last row and column."""
    w, h = 540, 70
    t = np.zeros((h, w), np.uint8)
    # (1) promoted run in tile (0,0) ends at (31,255); a weak pixel at (32,256) in tile (1,1) touches it only diagonally
    t[31, 250:256] = 128
    t[31, 250] = 255
    t[32, 256] = 128
    t[33, 257] = 128
    # (2) a closed blob inside tile (0,1) (dead) and one with a strong pixel (promoted)
    t[10:13, 300:303] = 128
    t[20:23, 300:303] = 128
    t[21, 301] = 255
    # (3) chain through tiles (1,0) -> (1,1) -> (0,1) -> ... seeded at its far end in tile (0,1)
    t[40, 200:300] = 128          # crosses x = 256 in the tile row below
    t[33:41, 299] = 128           # up towards the tile border row 32
    t[28:33, 298] = 128           # across y = 32, shifted by one column (diagonal contact)
    t[28, 298:330] = 128
    t[28, 330] = 255
    # (4) an open component that reaches two tiles but no strong pixel: dies
    t[50, 240:270] = 128
    # (5) last row / last column
    t[h - 1, 400:430] = 128
    t[h - 1, 400] = 255
    t[60:h, w - 1] = 128
    t[60, w - 1] = 255
    t[5:9, w - 1] = 128            # no seed: dies
    want = O.hysteresis(t)
    assert want[32, 256] == 255 and want[33, 257] == 255 and want[11, 301] == 0 and want[21, 300] == 255
    assert want[40, 200] == 255 and want[50, 250] == 0 and want[h - 1, 429] == 255 and want[h - 1, w - 1] == 255 and want[6, w - 1] == 0
    for spread in (8, 4):
        edges, bits = E.hysteresis(O.thresh_to_map2(t), w, spread=spread)
        assert np.array_equal(edges[0], want)
        assert np.array_equal(bits[0], O.edges_to_bits(want))
    # the transposed picture: the same contacts across the other kind of tile border
    tt = np.ascontiguousarray(t.T)
    edges, _ = E.hysteresis(O.thresh_to_map2(tt), h)
    assert np.array_equal(edges[0], O.hysteresis(tt))
