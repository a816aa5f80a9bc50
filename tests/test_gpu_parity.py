"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, the committed golden outputs of the
reference's kernels, and -- live -- the reference's own kernels compiled unmodified (oracle/_ref/libcvpref.so)."""
import glob
import os

import numpy as np
import pytest

import oracle_py as O
import cudacam_b200 as cb
from cudacam_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_FILES = sorted(glob.glob(os.path.join(GOLD, "frame_*.npz")))


def _check_all(c, f, r):
    c.run(f, cb.CannyStage.HYSTER)
    assert np.array_equal(c.edges(), r["edges"]), "edge map"
    assert np.array_equal(c.view(), r["edges"])
    assert np.array_equal(c.map2(), O.thresh_to_map2(r["thresh"])), "2-bit map"
    assert np.array_equal(c.bits(), O.edges_to_bits(r["edges"])), "bit plane"
    assert np.array_equal(c.mono(), r["mono"])
    assert np.array_equal(c.blur(), r["blur"])
    assert np.array_equal(c.gradient().view(np.uint32), r["grad"].view(np.uint32))
    assert np.array_equal(c.nms(), r["nms"])
    assert np.array_equal(c.thresh(), r["thresh"])


@pytest.mark.parametrize("impl", [0, 1], ids=["march", "tile"])
@pytest.mark.parametrize("kind,w,h,seed,lo,hi", [
    ("scene", 1280, 720, 0xC0FFEE, 10, 40),      # BASELINE config 1
    ("scene", 1280, 720, 0xC0FFEE, 17, 43),      # thresholds of the reference's screenshot
    ("noise", 641, 363, 2, 10, 40), ("steps", 800, 600, 3, 10, 40), ("steps", 333, 222, 4, 3, 200),
    ("scene", 31, 33, 6, 10, 40), ("noise", 1, 1, 1, 10, 40), ("scene", 16, 1, 2, 10, 40), ("scene", 5, 300, 2, 10, 40),
    ("scene", 1920, 1080, 5, 10, 40), ("steps", 1920, 1080, 8, 10, 40), ("noise", 1000, 700, 8, 17, 43), ("scene", 248, 2000, 4, 10, 40),
    ("scene", 1283, 721, 9, 10, 40), ("noise", 250, 129, 3, 10, 40), ("scene", 3840, 2160, 11, 10, 40),
])
def test_frame_vs_oracle(impl, kind, w, h, seed, lo, hi):
    f = synth.frame(kind, seed, w, h)
    r = O.canny(f, lo, hi)
    with cb.CannyEdge(w, h) as c:
        c.set_option("stencil_impl", impl)
        c.setHighThreshold(hi)
        c.setLowThreshold(lo)
        _check_all(c, f, r)


@pytest.mark.parametrize("path", FRAME_FILES, ids=[os.path.basename(p) for p in FRAME_FILES])
def test_frame_vs_golden(path):
    from test_oracle import parse
    kind, w, h, seed, lo, hi = parse(path)
    g = np.load(path)
    f = synth.frame(kind, seed, w, h)
    with cb.CannyEdge(w, h) as c:
        c.setHighThreshold(hi)
        c.setLowThreshold(lo)
        for stage in range(6):
            c.run(f, stage)
            assert np.array_equal(c.view(), g[f"pbo{stage}"]), f"stage {stage} view"
        assert np.array_equal(c.edges(), g["hyster"])
        assert np.array_equal(c.gradient().view(np.uint32), g["grad"].view(np.uint32))


def test_live_reference_kernels_720p():
    """BASELINE config 1: bit-exact to the reference's own cvp CUDA pipeline, run side by side on this GPU."""
    w, h = 1280, 720
    f = synth.frame("scene", 0xC0FFEE, w, h)
    ref = O.CvpRef(w, h)
    ref.run(f, 5)
    it, flag, _ = ref.info()
    assert flag == 0 and it < 100
    with cb.CannyEdge(w, h) as c:
        c.run(f)
        assert np.array_equal(c.edges(), ref.get("hyster"))
        assert np.array_equal(c.thresh(), ref.get("thresh"))
        assert np.array_equal(c.nms(), ref.get("nms"))
        assert np.array_equal(c.blur(), ref.get("blur"))
        assert np.array_equal(c.mono(), ref.get("mono"))
        assert np.array_equal(c.gradient().view(np.uint32), ref.get("grad").view(np.uint32))
    ref.close()


def test_hysteresis_worst_cases_and_launch_count():
    """On-device hysteresis against the fixpoint oracle, incl. a 1910-px weak line seeded at one end (65
    reference launches) and a dense random map; the launch count is fixed: 1 stencil + 3 hysteresis, whatever the chain
    length (the reference relaunches until a flag stays clear, capped at 100: cannyEdgeH.cu:297-338)."""
    from cudacam_b200 import _lib
    w, h = 1920, 64
    f = synth.frame("scene", 77, w, h)
    with cb.CannyEdge(w, h) as c:
        c.run(f)
        n0 = _lib.lib.b2c_launch_count(c._h)
        c.run(f)
        assert _lib.lib.b2c_launch_count(c._h) - n0 == 4
        assert np.array_equal(c.edges(), O.canny(f)["edges"])
    for kind, ww, hh, seed in (("noise", 640, 360, 5), ("steps", 800, 600, 6), ("scene", 3840, 2160, 9)):
        f = synth.frame(kind, seed, ww, hh)
        with cb.CannyEdge(ww, hh) as c:
            c.run(f)
            th = c.thresh()
            assert np.array_equal(c.edges(), O.hysteresis(th)), (kind, ww, hh)


def test_buffer_state_rules():
    """EDGES / BITS exist only after a run that reached HYSTER (no stale maps from an earlier frame); stage buffers of a
    device-resident run are rebuilt from the caller's input only while the handle still knows it."""
    w, h = 320, 200
    f, g = synth.frame("scene", 1, w, h), synth.frame("scene", 2, w, h)
    with cb.CannyEdge(w, h) as c:
        c.run(f)
        assert np.array_equal(c.edges(), O.canny(f)["edges"])
        c.run(g, cb.CannyStage.THRESH)
        assert np.array_equal(c.thresh(), O.canny(g)["thresh"])
        with pytest.raises(cb.B2cError):
            c.edges()
        with pytest.raises(cb.B2cError):
            c.bits()
        c.run(g)
        assert np.array_equal(c.edges(), O.canny(g)["edges"])


def test_strided_input_and_threshold_api():
    w, h = 300, 200
    big = np.zeros((h, 1024), np.uint8)
    f = synth.frame("scene", 9, w, h)
    big[:, :w * 3] = f.reshape(h, -1)
    view = np.lib.stride_tricks.as_strided(big, (h, w, 3), (1024, 3, 1))
    with cb.CannyEdge(w, h) as c:
        assert (c.getLowThreshold(), c.getHighThreshold()) == (10, 40)   # cannyEdgeH.cu:22-23
        c.setLowThreshold(200)
        assert c.getLowThreshold() == 40                                  # clamped: cannyEdgeH.hpp:25
        c.setHighThreshold(5)
        assert c.getHighThreshold() == 40
        c.setLowThreshold(10)
        assert c.isKernelProfilingEnabled()                               # default ON: cannyEdgeH.cu:24
        c.run(view)
        assert np.array_equal(c.edges(), O.canny(f)["edges"])
        t = c.lastTimings()   # event marks upload | stencil | hysteresis | output tile the total
        assert all(t[k] > 0 for k in ("upload", "stencil", "hysteresis", "output", "total")) and t["rounds"] == 1
        assert abs(t["upload"] + t["stencil"] + t["hysteresis"] + t["output"] - t["total"]) <= 0.02 * t["total"] + 0.005
        c.enableKernelProfiling(False)
        c.run(view)
        with pytest.raises(cb.B2cError):   # no marks were recorded for that run
            c.lastTimings()
        c.enableKernelProfiling(True)
        with pytest.raises(cb.B2cError):
            c.run(np.zeros((h + 1, w, 3), np.uint8))


def test_copy_view_is_the_pbo_copy():
    """b2c_copy_view = the D2D copy of cannyEdgeH.cu:188-207 into a caller-owned device buffer (the mapped PBO), for the
    edge map and for the saturated gradient view, with a pitch of its own."""
    import torch
    w, h = 333, 222
    f = synth.frame("steps", 4, w, h)
    r = O.canny(f)
    with cb.CannyEdge(w, h) as c:
        for stage, want in ((cb.CannyStage.HYSTER, r["edges"]), (cb.CannyStage.GRADIENT, O.float2uchar(r["grad"])), (cb.CannyStage.GAUSSIAN, r["blur"])):
            c.run(f, stage)
            for pitch in (0, 352):
                dst = torch.full((h, pitch or w), 7, dtype=torch.uint8, device="cuda")
                c.copy_view(dst.data_ptr(), pitch)
                c.sync()
                got = dst.cpu().numpy()
                assert np.array_equal(got[:, :w], want) and np.array_equal(got[:, :w], c.view())
                assert (got[:, w:] == 7).all()
        with pytest.raises(cb.B2cError):
            c.copy_view(dst.data_ptr(), w - 1)


def test_cvpipeline_surface():
    w, h = 160, 120
    f = synth.frame("scene", 3, w, h)
    p = cb.CvPipeline(0, w, h, 3)
    assert p.process(f, cb.CannyStage.HYSTER) is True
    assert np.array_equal(p.output(), O.canny(f)["edges"])
    assert p.process(np.zeros((0, 0, 3), np.uint8), cb.CannyStage.HYSTER) is False
    assert p.process(None, cb.CannyStage.HYSTER) is False
    assert p.process(f, cb.CannyStage.GRADIENT) is True
    assert np.array_equal(p.output(), O.float2uchar(O.canny(f)["grad"]))


def test_batch_host_pipeline_and_bits():
    w, h, n = 320, 180, 11
    frames = synth.batch("scene", n, w, h)
    want = np.stack([O.canny(frames[i])["edges"] for i in range(n)])
    with cb.CannyEdge(w, h, max_batch=4) as c:
        got = c.run_batch(frames)
        assert np.array_equal(got, want)
        bits = c.run_batch(frames, packed_bits=True)
        assert np.array_equal(bits, np.stack([O.edges_to_bits(want[i]) for i in range(n)]))
    with cb.CannyEdge(w, h, max_batch=1) as c:
        assert np.array_equal(c.run_batch(frames[:3]), want[:3])


def test_full_size_properties_1080p_batch():
    """BASELINE config 2 size (subset of the batch), checked through size-independent properties:
    strong pixels survive, non-candidates never appear, re-running hysteresis on the result is idempotent,
    and every frame equals the single-frame path."""
    w, h, n = 1920, 1080, 8
    frames = synth.batch("scene", n, w, h, distinct=4)
    with cb.CannyEdge(w, h, max_batch=8) as c:
        got = c.run_batch(frames)
    with cb.CannyEdge(w, h) as c1:
        for i in (0, 3, 7):
            c1.run(frames[i])
            assert np.array_equal(c1.edges(), got[i])
            th = c1.thresh()
            assert np.all(got[i][th == 255] == 255) and np.all(got[i][th == 0] == 0)
            t2 = np.where(got[i] == 255, 255, np.where(th == 128, 128, 0)).astype(np.uint8)
            assert np.array_equal(O.hysteresis(t2), got[i])
    assert np.array_equal(got[0], got[4])   # frames 0 and 4 are the same picture


@pytest.mark.parametrize("impl", [0, 1], ids=["march", "tile"])
@pytest.mark.parametrize("ch", [1, 4])
def test_other_input_formats(ch, impl):
    """GRAY8 and BGRA8 input: single frame with all accessors, and the batch pipeline."""
    w, h = 648, 360
    f = synth.frame("scene", 21, w, h)
    if ch == 1:
        g = np.ascontiguousarray(f[:, :, 1:2])
    else:
        g = np.ascontiguousarray(np.concatenate([f, np.random.default_rng(1).integers(0, 256, (h, w, 1), dtype=np.uint8)], axis=2))
    r = O.canny(g)
    if ch == 4:
        assert np.array_equal(r["edges"], O.canny(f)["edges"])   # alpha is ignored
    with cb.CannyEdge(w, h, channels=ch, max_batch=3) as c:
        c.set_option("stencil_impl", impl)
        _check_all(c, g, r)
        got = c.run_batch(np.stack([g, g, g]))
        assert np.array_equal(got[1], r["edges"])


def test_nv12_luma_plane():
    """NV12 camera frames: the luma plane is the gray picture -- the GRAY8 front end on the Y plane with
    the surface's pitch; the interleaved chroma plane behind it is never read."""
    w, h, pitch = 1000, 562, 1024
    f = synth.frame("scene", 17, w, h)
    y = ((f[:, :, 0].astype(np.uint32) * 7 + f[:, :, 1].astype(np.uint32) * 38 + f[:, :, 2].astype(np.uint32) * 19) >> 6).astype(np.uint8)
    nv12 = np.full((h + h // 2, pitch), 0x5A, np.uint8)   # Y rows, then the UV rows (garbage here)
    nv12[:h, :w] = y
    luma = np.lib.stride_tricks.as_strided(nv12, (h, w, 1), (pitch, 1, 1))
    r = O.canny(np.ascontiguousarray(y[:, :, None]))
    with cb.CannyEdge(w, h, channels=1) as c:
        c.run(luma)
        assert np.array_equal(c.edges(), r["edges"]) and np.array_equal(c.mono(), y)
    # and it is what the BGR path computes from the same picture (gray = (7B + 38G + 19R) >> 6, cannyEdgeD.cu:14-19)
    assert np.array_equal(r["edges"], O.canny(f)["edges"])


def test_planar_bgr_input():
    """Planar BGR8: three planes instead of interleaved pixels give the same result, through the
    single-frame call with all accessors, a strided surface and the batch pipeline."""
    w, h = 500, 281
    f = synth.frame("scene", 33, w, h)
    r = O.canny(f)
    planes = np.ascontiguousarray(f.transpose(2, 0, 1))
    with cb.CannyEdge(w, h, planar=True, max_batch=4) as c:
        c.run(planes)
        assert np.array_equal(c.edges(), r["edges"]) and np.array_equal(c.mono(), r["mono"]) and np.array_equal(c.nms(), r["nms"])
        big = np.zeros((3 * h, 512), np.uint8)
        big[:, :w] = planes.reshape(3 * h, w)
        c.run(np.lib.stride_tricks.as_strided(big, (3, h, w), (512 * h, 512, 1)), cb.CannyStage.THRESH)
        assert np.array_equal(c.thresh(), r["thresh"])
        got = c.run_batch(np.stack([planes] * 5))
        assert all(np.array_equal(got[i], r["edges"]) for i in range(5))
    with pytest.raises(ValueError):
        cb.CannyEdge(w, h, channels=1, planar=True)
