"""Row-band sharding of ONE large image over the ranks of a torch.distributed group (BASELINE config 5).

No reference counterpart: the reference is single-GPU.  Each rank owns a contiguous band of rows.
The fused stencil needs 4 rows of INPUT beyond each interior seam (2 Gaussian + 1 Sobel + 1 NMS), exchanged once with
the neighbours (NCCL send/recv over NVLink on GPUs); the reference's zero padding applies only at the true image
border.  Hysteresis: every rank resolves its band on the device (union-find kernel), then the first/last row of its
edge bit-plane goes to the neighbours' ghost rows; ranks whose ghost rows gained bits re-resolve with those rows as
extra seeds; the loop ends when an all-reduce (MAX) of "my ghost rows changed" is 0.  The result is bit-identical to
the unsharded run because both compute the same fixpoint.

torch.distributed is plumbing here (rendezvous, send/recv, all_reduce); all pixel work is behind `backend`:
`CudaBandBackend` (the C ABI of libb200canny.so) on GPUs, and an emulator-based stand-in in tests (gloo, CPU).
"""
import ctypes as C

import numpy as np

from . import _lib

HALO = 4


def band_rows(height, world, rank):
    """Contiguous split: the first height % world bands get one extra row.  Returns (y0, rows)."""
    base, extra = divmod(height, world)
    rows = base + (1 if rank < extra else 0)
    y0 = rank * base + min(rank, extra)
    return y0, rows


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view device memory owned by the C library."""

    def __init__(self, ptr, n, typestr):
        shape = tuple(n) if isinstance(n, (tuple, list)) else (n,)
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(int(ptr), False), version=3)


class CudaBandBackend:
    """Band compute through the C ABI (b2c_create_band / b2c_band_*)."""

    def __init__(self, width, rows, y0, height_global, device):
        import torch
        self.torch = torch
        self.width, self.rows, self.y0, self.height_global, self.device = width, rows, y0, height_global, device
        h = C.c_void_p()
        _lib.check(_lib.lib.b2c_create_band(C.byref(h), device, width, rows, y0, height_global), what="b2c_create_band")
        self._h = h
        # the input buffer (HALO rows above, the band, HALO rows below) is owned by the C library so that the neighbour
        # ranks can map it (CUDA IPC) and store their halo rows straight into it
        p, st = C.c_void_p(), C.c_size_t()
        _lib.check(_lib.lib.b2c_band_input(h, C.byref(p), C.byref(st)), h, "b2c_band_input")
        self.row_stride = st.value
        with torch.cuda.device(device):
            self.buf = torch.as_tensor(_DevArray(p.value, (rows + 2 * HALO, self.row_stride), "|u1"), device=f"cuda:{device}")
        self.wpr = (width + 31) // 32
        self._views = {}

    def close(self):
        if self._h:
            _lib.lib.b2c_destroy(self._h)
            self._h = None

    def _stream(self):
        # torch's current stream on this band's device; handle 0 is the legacy default stream, which the C ABI
        # spells cudaStreamLegacy (0x1) because 0 means "the handle's own stream" there
        return self.torch.cuda.current_stream(self.device).cuda_stream or 1

    def input_rows(self, r0, r1):
        """Tensor view of buffer rows [r0, r1) (row HALO is band row 0)."""
        return self.buf[r0:r1]

    def load(self, band_host):
        """band_host: (rows, w, 3) uint8 -> device band rows."""
        t = self.torch.from_numpy(np.ascontiguousarray(band_host).reshape(self.rows, self.width * 3))
        self.buf[HALO:HALO + self.rows, :self.width * 3].copy_(t, non_blocking=False)

    def stencil(self):
        ptr = self.buf.data_ptr() + HALO * self.row_stride
        _lib.check(_lib.lib.b2c_band_stencil(self._h, ptr, self.row_stride, self._stream()), self._h, "b2c_band_stencil")

    def hysteresis(self, first, write_edges):
        """write_edges: False = bit plane only, True = also the u8 map, "only" = just expand the final bit plane."""
        we = 2 if write_edges == "only" else 1 if write_edges else 0
        _lib.check(_lib.lib.b2c_band_hysteresis(self._h, 1 if first else 0, we, None, self._stream()), self._h, "b2c_band_hysteresis")

    def enable_p2p(self, dist, rank, world, group=None):
        """Maps the other ranks' mailboxes (CUDA IPC) so that the cross-band rounds run on the devices (NVLink peer
        stores, device-side convergence) instead of through NCCL + host.  Ranks must be on one box."""
        h = (C.c_ubyte * 144)()
        _lib.check(_lib.lib.b2c_band_p2p_export(self._h, h), self._h, "b2c_band_p2p_export")
        mine = self.torch.tensor(list(h), dtype=self.torch.uint8, device=f"cuda:{self.device}")
        allh = self.torch.empty(144 * world, dtype=self.torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        buf = bytes(allh.cpu().tolist())
        _lib.check(_lib.lib.b2c_band_p2p_open(self._h, buf, world, rank), self._h, "b2c_band_p2p_open")
        self.p2p = True

    def halo_p2p(self):
        _lib.check(_lib.lib.b2c_band_p2p_halo(self._h, self._stream()), self._h, "b2c_band_p2p_halo")

    def converge(self, rounds_per_sync=16):
        n = C.c_int(0)
        _lib.check(_lib.lib.b2c_band_p2p_converge(self._h, rounds_per_sync, C.byref(n), self._stream()), self._h, "b2c_band_p2p_converge")
        return n.value

    def seeded(self):
        """int32[1] device tensor: 1 if the last re-entry call found a ghost pixel that seeded something new."""
        if "flag" not in self._views:
            p = C.c_void_p()
            _lib.check(_lib.lib.b2c_band_flag_ptr(self._h, C.byref(p)), self._h, "flag")
            self._views["flag"] = self.torch.as_tensor(_DevArray(p.value, 1, "<i4"), device=f"cuda:{self.device}")
        return self._views["flag"]

    def _view(self, kind, which):
        key = (kind, which)
        if key not in self._views:
            p, n = C.c_void_p(), C.c_int()
            f = _lib.lib.b2c_band_boundary_ptr if kind == "boundary" else _lib.lib.b2c_band_ghost_ptr
            _lib.check(f(self._h, which, C.byref(p), C.byref(n)), self._h, kind)
            self._views[key] = self.torch.as_tensor(_DevArray(p.value, n.value, "<i4"), device=f"cuda:{self.device}")
        return self._views[key]

    def boundary(self, which):
        """int32 tensor view of the first (0) / last (1) row of the edge bit-plane."""
        return self._view("boundary", which)

    def ghost(self, which):
        """int32 tensor view of the ghost row above (0) / below (1) the band."""
        return self._view("ghost", which)

    def sync(self):
        self.torch.cuda.synchronize(self.device)

    @property
    def launches(self):
        """Kernels launched by this band's handle so far."""
        return _lib.lib.b2c_launch_count(self._h)

    def edges(self):
        out = np.empty((self.rows, self.width), np.uint8)
        self.torch.cuda.synchronize(self.device)
        _lib.check(_lib.lib.b2c_download(self._h, _lib.BUF_EDGES, out.ctypes.data, out.strides[0]), self._h, "b2c_download")
        return out


class BandCanny:
    """Drives one rank's band: halo exchange, stencil, cross-band hysteresis to the global fixpoint."""

    def __init__(self, backend, rank, world, dist=None, group=None):
        self.b, self.rank, self.world, self.dist, self.group = backend, rank, world, dist, group
        self.rounds = 0

    # -- plumbing ------------------------------------------------------------------------------------------------
    def _exchange(self, send_up, recv_up, send_down, recv_down):
        """send_up goes to rank-1 (lands in its recv_down), send_down to rank+1 (its recv_up)."""
        if self.world == 1:
            return
        d = self.dist
        ops = []
        if self.rank > 0:
            ops += [d.P2POp(d.isend, send_up, self.rank - 1, self.group), d.P2POp(d.irecv, recv_up, self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            ops += [d.P2POp(d.isend, send_down, self.rank + 1, self.group), d.P2POp(d.irecv, recv_down, self.rank + 1, self.group)]
        for r in d.batch_isend_irecv(ops):
            r.wait()

    def exchange_input_halos(self):
        b, n = self.b, self.b.rows
        if self.world > 1 and getattr(b, "p2p", False):
            b.halo_p2p()   # peer stores into the neighbours' buffers + device-side arrival counters
            return
        # contiguous staging: rows of a strided buffer are contiguous blocks already (full-stride rows)
        self._exchange(b.input_rows(HALO, 2 * HALO), b.input_rows(0, HALO), b.input_rows(n, n + HALO), b.input_rows(n + HALO, n + 2 * HALO))

    def run(self):
        """Stencil + hysteresis of this band; returns the number of global hysteresis rounds used."""
        b, d = self.b, self.dist
        b.ghost(0).zero_()   # ghost rows = the image border's zero padding until a neighbour says otherwise
        b.ghost(1).zero_()
        self.exchange_input_halos()
        b.stencil()
        # band-local fixpoint (planes + union-find forest are kept).  Every resolve pass also writes the u8 edge map,
        # so there is no separate expansion pass once the rounds have converged.
        b.hysteresis(True, write_edges=True)
        rounds = 1
        if self.world > 1 and getattr(b, "p2p", False):
            rounds = max(1, b.converge() - 1)   # device-side rounds over NVLink peer memory (the last one finds nothing new)
        while self.world > 1 and not getattr(b, "p2p", False):
            # boundary rows of the edge bit-plane -> the neighbours' ghost rows; re-entry seeds the weak runs that
            # touch a strong ghost pixel and resolves their components; stop when no rank was seeded anything new.
            # (One all_gather of rows + flag per round instead of send/recv + all_reduce was measured SLOWER: the
            # extra small tensor ops on the host cost more than the second NCCL launch.)
            self._exchange(b.boundary(0), b.ghost(0), b.boundary(1), b.ghost(1))
            b.hysteresis(False, write_edges=True)
            flag = b.seeded().clone()
            d.all_reduce(flag, op=d.ReduceOp.MAX, group=self.group)
            if int(flag.item()) == 0:
                break
            rounds += 1
        self.rounds = rounds
        return rounds


def run_local(backends):
    """All bands driven from ONE process (bands on the same or on different GPUs; copies between devices go over
    NVLink peer-to-peer).  Same protocol as BandCanny.run without a process group: used for the 1-GPU data point of
    config 5, for the single-GPU tests of the band logic, and by hosts that prefer one thread for the whole box.
    Returns the number of global hysteresis rounds."""
    n = len(backends)
    for b in backends:
        b.ghost(0).zero_()
        b.ghost(1).zero_()
    for i in range(n - 1):
        up, dn = backends[i], backends[i + 1]
        dn.input_rows(0, HALO).copy_(up.input_rows(up.rows, up.rows + HALO), non_blocking=True)
        up.input_rows(up.rows + HALO, up.rows + 2 * HALO).copy_(dn.input_rows(HALO, 2 * HALO), non_blocking=True)
    for b in backends:
        b.sync()
    for b in backends:
        b.stencil()
    for b in backends:
        b.hysteresis(True, write_edges=True)
    rounds = 1
    while n > 1:
        for b in backends:
            b.sync()
        for i in range(n - 1):
            up, dn = backends[i], backends[i + 1]
            dn.ghost(0).copy_(up.boundary(1))
            up.ghost(1).copy_(dn.boundary(0))
        for b in backends:
            b.sync()
        for b in backends:
            b.hysteresis(False, write_edges=True)
        if not any(bool(b.seeded().item()) for b in backends):
            break
        rounds += 1
    for b in backends:
        b.sync()
    return rounds
