"""GPU tests of the two partitioned modes: row bands of one image (bit-identical to the unsharded
run) and frame-parallel handles on several devices.  The single-GPU cases run everywhere; the NCCL cases need >= 2
GPUs (`gpurun --gpus 2`) and skip otherwise."""
import os
import socket

import numpy as np
import pytest

import oracle_py as O
import cudacam_b200 as cb
from cudacam_b200 import _lib, bands, synth

pytestmark = pytest.mark.gpu


def _ndev():
    return _lib.lib.b2c_device_count()


def _unsharded(img):
    h, w, _ = img.shape
    with cb.CannyEdge(w, h) as c:
        c.run(img)
        return c.edges()


@pytest.mark.parametrize("kind,w,h,nb", [("scene", 1920, 1000, 3), ("steps", 640, 203, 5), ("noise", 328, 64, 2), ("scene", 4096, 4096, 4)])
def test_local_bands_equal_unsharded(kind, w, h, nb):
    """nb bands driven from one process (all on cuda:0, or spread over the devices present)."""
    img = synth.frame(kind, 11, w, h)
    want = _unsharded(img)
    if h <= 1100:
        assert np.array_equal(want, O.canny(img)["edges"])
    nd = max(1, _ndev())
    bes = []
    for r in range(nb):
        y0, rows = bands.band_rows(h, nb, r)
        b = bands.CudaBandBackend(w, rows, y0, h, device=r % nd)
        b.load(img[y0:y0 + rows])
        bes.append(b)
    assert bands.run_local(bes) == 1
    got = np.concatenate([b.edges() for b in bes])
    assert np.array_equal(got, want)
    # the same bands through the peer-to-peer kernels (halo push / seam push + wait), every band on its own stream
    bands.open_local(bes)
    for rep in range(2):
        assert bands.run_local(bes) == 1
        assert all(b.status()[1] == 0 for b in bes)
        assert np.array_equal(np.concatenate([b.edges() for b in bes]), want), rep
    for b in bes:
        b.close()
    if kind == "scene" and h <= 1100:
        # the north star's seam criterion (<= 0.1 % extra disagreement with cv::Canny at the band seams): sharding adds 0
        import test_cv2_disagreement as T
        assert abs(T.seam_excess_vs_unsharded(got, want, T.cv_canny(img), nb)) <= 0.1


def _thresh_bands(t, nb, p2p, force_global=False):
    """Hysteresis-only band run on a given thresholded map (u8 0/128/255) split into nb bands on cuda:0."""
    h, w = t.shape
    bes = []
    for r in range(nb):
        y0, rows = bands.band_rows(h, nb, r)
        b = bands.CudaBandBackend(w, rows, y0, h, device=0)
        b.load_thresh(t[y0:y0 + rows])
        if force_global:
            _lib.check(_lib.lib.b2c_set_option(b._h, b"seam_force_global", 1), b._h, "set_option")
        bes.append(b)
    if p2p:
        bands.open_local(bes)
    bands.run_local(bes, stencil=False)
    got = np.concatenate([b.edges() for b in bes])
    err = [b.status()[1] for b in bes]
    for b in bes:
        b.close()
    assert not any(err)
    return got


@pytest.mark.parametrize("p2p", [False, True], ids=["collective", "p2p"])
def test_seam_solve_snake_and_clutter_on_device(p2p):
    """The real seam kernels on one GPU: a weak snake that crosses every seam ~20 times (one exchange must do), and
    random clutter over 8 bands of unequal height; wide enough for several words per row."""
    from test_bands_gloo import _snake_map
    t = _snake_map(300, 67)
    assert np.array_equal(_thresh_bands(t, 5, p2p), O.hysteresis(t))
    rng = np.random.default_rng(3)
    t = np.where(rng.random((203, 1500)) < 0.4, 128, 0).astype(np.uint8)
    t[rng.random(t.shape) < 0.001] = 255
    assert np.array_equal(_thresh_bands(t, 8, p2p), O.hysteresis(t))
    assert np.array_equal(_thresh_bands(t, 8, p2p, force_global=True), O.hysteresis(t))   # global-memory hash / forest


def test_giga_bands_hash_equals_golden():
    """BASELINE config 5 on one GPU: 8 bands of the 16384^2 mosaic through the peer-to-peer kernels; sha256 of the
    assembled edge map against the committed hash of the oracle's output (tests/golden/giga_sha256.json)."""
    import hashlib, json
    w = h = 16384
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "giga_sha256.json")))
    nb = 8
    bes = []
    for r in range(nb):
        y0, rows = bands.band_rows(h, nb, r)
        b = bands.CudaBandBackend(w, rows, y0, h, device=0)
        b.load(synth.giga_rows(y0, y0 + rows, w, h))
        bes.append(b)
    bands.open_local(bes)
    bands.run_local(bes)
    sha = hashlib.sha256()
    for b in bes:
        sha.update(b.edges().tobytes())
        b.close()
    assert sha.hexdigest() == gold["16384x16384"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, w, h, q, p2p):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        img = synth.frame("scene", 13, w, h)
        y0, rows = bands.band_rows(h, world, rank)
        be = bands.CudaBandBackend(w, rows, y0, h, device=rank)
        be.load(img[y0:y0 + rows])
        if p2p and not be.enable_p2p(dist, rank, world):   # (collective answer: the same on every rank)
            q.put((rank, 0, 0, -2, "CUDA IPC / peer access not available on this box"))
            be.close()
            return
        bc = bands.BandCanny(be, rank, world, dist)
        bc.run()
        rounds = bc.run()   # twice: the second run re-uses planes, forest, mailboxes and run counters
        err = be.status()[1]
        q.put((rank, y0, rows, -1 if err else rounds, be.edges()))
        be.close()
    except BaseException as e:
        q.put((rank, 0, 0, -1, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True], ids=["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_bands_equal_unsharded(world, p2p):
    if _ndev() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    w, h = 2048, 1536
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, w, h, q, p2p)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(world)]
    if any(r[3] == -2 for r in res):
        for p in procs:
            p.join(300)
        pytest.skip(res[0][4])
    assert all(r[3] == 1 for r in res), [r[:4] for r in res]
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = np.zeros((h, w), np.uint8)
    for rank, y0, rows, rounds, e in res:
        got[y0:y0 + rows] = e
    assert np.array_equal(got, _unsharded(synth.frame("scene", 13, w, h)))


def test_frame_parallel_handles_on_every_device():
    """One independent handle per GPU, no shared state (frame batches split with no collective)."""
    w, h, n = 640, 360, 6
    frames = synth.batch("scene", n, w, h)
    want = np.stack([O.canny(frames[i])["edges"] for i in range(n)])
    for dev in range(max(1, _ndev())):
        with cb.CannyEdge(w, h, device=dev, max_batch=4) as c:
            assert np.array_equal(c.run_batch(frames), want), dev
