"""Row-band sharding on the CPU: world_size 2 and 3 over gloo, kernels through the emulator (tests/emu), compared
with the oracle on the whole image.  Covers the host logic of cudacam_b200/bands.py (halo exchange of 4 input rows,
all-gather of the seam records) and the seam kernels themselves (k_seam_publish / k_seam_solve + the resolve pass)
without a GPU; the NCCL/GPU twin of this test is tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emu_py as E
import oracle_py as O
from cudacam_b200 import bands, synth


class EmuBandBackend:
    """Same interface as bands.CudaBandBackend; pixel work by the emulated kernels (test-only)."""

    def __init__(self, width, rows, y0, height_global, thresh_override=None, ucap=0, force_global=False):
        self.width, self.rows, self.y0, self.height_global = width, rows, y0, height_global
        self.row_stride = (width * 3 + 15) // 16 * 16
        self.buf = torch.zeros((rows + 2 * bands.HALO, self.row_stride), dtype=torch.uint8)
        self._band = E.Band(width, rows, ucap, force_global)
        self.seam_bytes = self._band.seam_words * 4
        self._map2 = None
        self._all = None
        self._override = thresh_override   # (rows, w) u8 {0,128,255}: skips the stencil (hysteresis-only tests)

    def input_rows(self, r0, r1):
        return self.buf[r0:r1]

    def load(self, band_host):
        t = torch.from_numpy(np.ascontiguousarray(band_host).reshape(self.rows, self.width * 3))
        self.buf[bands.HALO:bands.HALO + self.rows, :self.width * 3] = t

    def stencil(self):
        if self._override is not None:
            self._map2 = O.thresh_to_map2(self._override)
            return
        a = self.buf.numpy()
        if self.width >= 8:
            self._map2 = E.stencil_raw(a, bands.HALO, self.width, self.rows, impl=120, y0=self.y0, h_glob=self.height_global)
        else:
            f = np.ascontiguousarray(a[:, :self.width * 3]).reshape(a.shape[0], self.width, 3)
            self._map2 = E.stencil(f, impl=1, y0=self.y0, h_glob=self.height_global, rows=self.rows, row0=bands.HALO)["map2"][0]

    def hysteresis(self):
        self._band.hysteresis(self._map2)

    def seam_record(self):
        return torch.from_numpy(self._band.record.view(np.uint8))

    def gather_buffer(self, world):
        if self._all is None or self._all.numel() != world * self.seam_bytes:
            self._all = torch.empty(world * self.seam_bytes, dtype=torch.uint8)
        return self._all

    def seam_solve(self, all_records, world, rank):
        self.promoted = self._band.solve(all_records.numpy().view(np.uint32), world, rank)

    def sync(self):
        pass

    def edges(self):
        return self._band.edges()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, w, h, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y0, rows = bands.band_rows(h, world, rank)
        if mode == "image":
            img = synth.frame("scene", 31, w, h)
            be = EmuBandBackend(w, rows, y0, h)
            be.load(img[y0:y0 + rows])
        else:
            t = _snake_map(w, h)
            be = EmuBandBackend(w, rows, y0, h, thresh_override=t[y0:y0 + rows])
        bc = bands.BandCanny(be, rank, world, dist)
        rounds = bc.run()
        q.put((rank, y0, rows, rounds, be.edges().copy()))
    except BaseException as e:   # the parent must not wait forever for a result
        q.put((rank, 0, 0, -1, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def _snake_map(w, h):
    """A weak chain that crosses every seam several times, seeded by ONE strong pixel in the top band, plus weak
    islands that must vanish: needs several global rounds."""
    t = np.zeros((h, w), np.uint8)
    x = 2
    up = False
    while x + 6 < w:
        t[2:h - 2, x] = 128           # full-height vertical stroke
        if up:
            t[2, x:x + 6] = 128       # connect at the top
        else:
            t[h - 3, x:x + 6] = 128   # connect at the bottom
        up = not up
        x += 6
    t[2, 2] = 255
    t[h // 2, w - 2] = 128            # isolated
    return t


def _run(world, mode, w, h):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(world)]
    assert all(r[3] >= 0 for r in res), res
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    out = np.zeros((h, w), np.uint8)
    for rank, y0, rows, rounds, e in res:
        out[y0:y0 + rows] = e
    return out, max(r[3] for r in res)


def test_band_rows_partition():
    for h, world in ((16384, 8), (1080, 7), (10, 3), (5, 5)):
        spans = [bands.band_rows(h, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(s[1] for s in spans) == h
        for a, b in zip(spans, spans[1:]):
            assert a[0] + a[1] == b[0]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_image_equals_unsharded(world):
    w, h = 96, 60
    got, rounds = _run(world, "image", w, h)
    want = O.canny(synth.frame("scene", 31, w, h))["edges"]
    assert np.array_equal(got, want)
    assert rounds >= 1


def test_cross_band_hysteresis_one_exchange():
    """The snake crosses the seam ~10 times; the earlier protocol needed one round per crossing, the seam solve one."""
    w, h = 64, 36
    got, exchanges = _run(2, "snake", w, h)
    want = O.hysteresis(_snake_map(w, h))
    assert np.array_equal(got, want)
    assert want[h - 3, w - 10] == 255 and got[h // 2, w - 2] == 0
    assert exchanges == 1


def _local(t, world, ucap=0, force_global=False):
    h, w = t.shape
    bes = []
    for r in range(world):
        y0, rows = bands.band_rows(h, world, r)
        bes.append(EmuBandBackend(w, rows, y0, h, thresh_override=t[y0:y0 + rows], ucap=ucap, force_global=force_global))
    n = bands.run_local(bes)
    return np.concatenate([b.edges() for b in bes]), n


def test_run_local_single_process_bands():
    """The single-process driver (bands.run_local) follows the same protocol: 3 bands of unequal height."""
    t = _snake_map(64, 37)
    got, n = _local(t, 3)
    assert np.array_equal(got, O.hysteresis(t)) and n == 1


@pytest.mark.parametrize("force_global", [False, True], ids=["smem", "gmem"])
@pytest.mark.parametrize("world", [2, 4, 7])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_seam_solve_random_maps(world, seed, force_global):
    """Random weak clutter with few strong pixels: many components thread through several bands, some reach an edge
    only through a chain of other bands' unresolved components.  Widths around the 32-bit word boundaries."""
    rng = np.random.default_rng(seed)
    w = (33, 64, 131)[seed]
    h = world * 4 + int(rng.integers(0, 9))
    r = rng.random((h, w))
    t = np.where(r < 0.42, 128, 0).astype(np.uint8)
    t[rng.random((h, w)) < 0.004] = 255
    got, _ = _local(t, world, force_global=force_global)
    assert np.array_equal(got, O.hysteresis(t))


def test_unresolved_word_list_overflow_falls_back_to_full_pass():
    """A list capacity of 3 words: the pass after the solve must visit the whole band instead."""
    rng = np.random.default_rng(9)
    t = np.where(rng.random((40, 200)) < 0.4, 128, 0).astype(np.uint8)
    t[rng.random(t.shape) < 0.003] = 255
    got, _ = _local(t, 3, ucap=3)
    assert np.array_equal(got, O.hysteresis(t))


def test_seam_solve_minimum_band_height():
    """Bands of 4 rows (the minimum): first and last row of a band nearly coincide; all-weak image with one strong
    pixel in the last band promotes everything, without it nothing."""
    w, h, world = 40, 16, 4
    t = np.full((h, w), 128, np.uint8)
    got, _ = _local(t, world)
    assert got.sum() == 0
    t[h - 1, w - 1] = 255
    got, _ = _local(t, world)
    assert (got == 255).all()


def test_band_handles_are_reusable():
    """A second frame through the same band objects (hash / forest / control words are re-initialised)."""
    w, h, world = 64, 37, 3
    bes = []
    for r in range(world):
        y0, rows = bands.band_rows(h, world, r)
        bes.append(EmuBandBackend(w, rows, y0, h))
    for t in (_snake_map(w, h), np.zeros((h, w), np.uint8), _snake_map(w, h)[::-1].copy()):
        for r, b in enumerate(bes):
            b._override = t[b.y0:b.y0 + b.rows]
        bands.run_local(bes)
        assert np.array_equal(np.concatenate([b.edges() for b in bes]), O.hysteresis(t))
