"""Row-band sharding of ONE large image over the ranks of a torch.distributed group (BASELINE config 5).

No reference counterpart: the reference is single-GPU.  Each rank owns a contiguous band of rows.
The fused stencil needs 4 rows of INPUT beyond each interior seam (2 Gaussian + 1 Sobel + 1 NMS), exchanged once with
the neighbours; the reference's zero padding applies only at the true image border.  Hysteresis: every rank builds the
union-find forest of its band, writes its seam record (edge / unresolved-weak bit rows of its first and last row + a
component label per unresolved run) and resolves the band while the record travels; then ONE exchange: the records
are all-gathered, and every rank solves the same small connected-components problem over all seams and promotes its own
components that reach an edge pixel of any band.  The result is bit-identical to the unsharded run
because both compute the same fixpoint.

Transports: peer memory (ranks of one box: halo rows and records are plain stores into the peers' mapped buffers) or
torch.distributed collectives (NCCL on GPUs, gloo in the CPU tests).  torch.distributed is plumbing here; all pixel
work is behind `backend`: `CudaBandBackend` (the C ABI of libb200canny.so) on GPUs, an emulator-based stand-in in
tests.  The C++ twin of this driver is b2c::BandRunner (include/b200canny.hpp).
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
        n = C.c_size_t()
        _lib.check(_lib.lib.b2c_band_seam_bytes(h, C.byref(n)), h, "b2c_band_seam_bytes")
        self.seam_bytes = n.value
        self.p2p = False
        self.own_stream = False   # launch on the handle's own stream instead of torch's current one (see open_local)
        self._all = None

    def close(self):
        if self._h:
            _lib.lib.b2c_destroy(self._h)
            self._h = None

    def _stream(self):
        # torch's current stream on this band's device; handle 0 is the legacy default stream, which the C ABI
        # spells cudaStreamLegacy (0x1) because 0 means "the handle's own stream" there
        if self.own_stream:
            return None
        return self.torch.cuda.current_stream(self.device).cuda_stream or 1

    def input_rows(self, r0, r1):
        """Tensor view of buffer rows [r0, r1) (row HALO is band row 0)."""
        return self.buf[r0:r1]

    def load(self, band_host):
        """band_host: (rows, w, 3) uint8 -> device band rows."""
        t = self.torch.from_numpy(np.ascontiguousarray(band_host).reshape(self.rows, self.width * 3))
        self.buf[HALO:HALO + self.rows, :self.width * 3].copy_(t, non_blocking=False)

    def load_thresh(self, thresh_host):
        """(rows, w) uint8 map of 0 / 128 / 255 in place of a stencil run (hysteresis on maps produced elsewhere)."""
        t = np.ascontiguousarray(thresh_host, np.uint8)
        assert t.shape == (self.rows, self.width)
        _lib.check(_lib.lib.b2c_load_thresh(self._h, t.ctypes.data, t.strides[0]), self._h, "b2c_load_thresh")

    def stencil(self):
        ptr = self.buf.data_ptr() + HALO * self.row_stride
        _lib.check(_lib.lib.b2c_band_stencil(self._h, ptr, self.row_stride, self._stream()), self._h, "b2c_band_stencil")

    def hysteresis(self):
        """Band-local fixpoint (planes and forest are kept for the seam solve); writes the u8 edge map and the band's
        seam record (with peer wiring: also into every rank's mailbox)."""
        _lib.check(_lib.lib.b2c_band_hysteresis(self._h, self._stream()), self._h, "b2c_band_hysteresis")

    # -- cross-band hysteresis: collective transport -------------------------------------------------------------
    def seam_record(self):
        """This band's seam record as a uint8 device tensor (owned by the handle, written by hysteresis())."""
        p = C.c_void_p()
        _lib.check(_lib.lib.b2c_band_seam_record(self._h, C.byref(p)), self._h, "b2c_band_seam_record")
        return self.torch.as_tensor(_DevArray(p.value, self.seam_bytes, "|u1"), device=f"cuda:{self.device}")

    def gather_buffer(self, world):
        if self._all is None or self._all.numel() != world * self.seam_bytes:
            self._all = self.torch.empty(world * self.seam_bytes, dtype=self.torch.uint8, device=f"cuda:{self.device}")
        return self._all

    def seam_solve(self, all_records, world, rank):
        _lib.check(_lib.lib.b2c_band_seam_solve(self._h, all_records.data_ptr(), world, rank, self._stream()), self._h, "b2c_band_seam_solve")

    # -- cross-band hysteresis and input halo: peer memory -------------------------------------------------------
    def enable_p2p(self, dist, rank, world, group=None):
        """Maps the other ranks' mailboxes and input buffers (CUDA IPC).  Ranks must be on one box.  Collective: every
        rank gets the same answer -- True, or False if ANY rank could not export or map (no CUDA IPC in this container,
        no peer access ...); the band then keeps the collective transport."""
        torch = self.torch
        dev = f"cuda:{self.device}"

        def all_ok(ok):
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(t.item())

        h = (C.c_ubyte * 144)()
        ok = _lib.lib.b2c_band_p2p_export(self._h, h) == 0
        mine = torch.tensor(list(h), dtype=torch.uint8, device=dev)
        allh = torch.empty(144 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        if not all_ok(ok):
            return False
        buf = bytes(allh.cpu().tolist())
        ok = _lib.lib.b2c_band_p2p_open(self._h, buf, world, rank) == 0
        self.p2p = all_ok(ok)
        return self.p2p

    def stencil_p2p(self, phase=0):
        """Halo exchange over peer memory hidden behind the stencil: only the CTAs next to a seam wait for the rows.
        phase: 0 = all (one band per process), 1 = push, 2 = stencil (see run_local)."""
        _lib.check(_lib.lib.b2c_band_p2p_stencil(self._h, self._stream(), phase), self._h, "b2c_band_p2p_stencil")

    def seam_p2p(self):
        _lib.check(_lib.lib.b2c_band_p2p_seam(self._h, self._stream()), self._h, "b2c_band_p2p_seam")

    def seam_phase_us(self):
        """Phase times of the last hysteresis + seam pass (needs set_phase_timing(True)): dict of microseconds."""
        names = ("tile_border", "publish_push", "resolve", "gap", "wait_solve", "list_pass")
        d = {n: _lib.lib.b2c_get_info(self._h, b"seam_phase_us%d" % k) for k, n in enumerate(names)}
        if self.p2p:
            d.update({n: _lib.lib.b2c_get_info(self._h, b"band_stencil_us%d" % k) for k, n in enumerate(("halo_push", "gap", "stencil"))})
        return d

    def set_phase_timing(self, on):
        _lib.check(_lib.lib.b2c_set_option(self._h, b"hyst_phase_timing", 1 if on else 0), self._h, "b2c_set_option")

    def status(self):
        """(weak runs promoted by the last solve, peer time-out flag) -- blocking."""
        n, e = C.c_int(0), C.c_int(0)
        _lib.check(_lib.lib.b2c_band_status(self._h, C.byref(n), C.byref(e)), self._h, "b2c_band_status")
        return n.value, e.value

    def sync(self):
        self.torch.cuda.synchronize(self.device)

    def mark(self):
        """An event recorded on the stream the band's work is launched on (torch's current stream)."""
        assert not self.own_stream
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.torch.cuda.current_stream(self.device))
        return e

    @property
    def launches(self):
        """Kernels launched by this band's handle so far."""
        return _lib.lib.b2c_launch_count(self._h)

    def edges(self):
        out = np.empty((self.rows, self.width), np.uint8)
        self.torch.cuda.synchronize(self.device)
        _lib.check(_lib.lib.b2c_download(self._h, _lib.BUF_EDGES, out.ctypes.data, out.strides[0]), self._h, "b2c_download")
        return out


def open_local(backends):
    """Peer wiring for CudaBandBackends that live in ONE process (one or several devices): every band gets the others'
    mailboxes and input buffers as plain pointers.  The bands then launch on their handles' OWN streams: the peer-to-
    peer kernels wait on the device for the other bands' stores, which on one shared stream would never be issued."""
    n = len(backends)
    arr = (C.c_void_p * n)(*[b._h for b in backends])
    for r, b in enumerate(backends):
        _lib.check(_lib.lib.b2c_band_p2p_open_local(b._h, arr, n, r), b._h, "b2c_band_p2p_open_local")
        b.p2p = True
        b.own_stream = True


class BandCanny:
    """Drives one rank's band: halo exchange, stencil, band-local hysteresis, one seam exchange."""

    def __init__(self, backend, rank, world, dist=None, group=None):
        self.b, self.rank, self.world, self.dist, self.group = backend, rank, world, dist, group
        self.exchanges = 0

    def _exchange(self, send_up, recv_up, send_down, recv_down):
        """send_up goes to rank-1 (lands in its recv_down), send_down to rank+1 (its recv_up)."""
        d = self.dist
        ops = []
        if self.rank > 0:
            ops += [d.P2POp(d.isend, send_up, self.rank - 1, self.group), d.P2POp(d.irecv, recv_up, self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            ops += [d.P2POp(d.isend, send_down, self.rank + 1, self.group), d.P2POp(d.irecv, recv_down, self.rank + 1, self.group)]
        for r in d.batch_isend_irecv(ops):
            r.wait()

    def exchange_input_halos(self):
        """Collective transport: the 4 input rows on either side of every seam by send / recv."""
        b, n = self.b, self.b.rows
        if self.world == 1:
            return
        self._exchange(b.input_rows(HALO, 2 * HALO), b.input_rows(0, HALO), b.input_rows(n, n + HALO), b.input_rows(n + HALO, n + 2 * HALO))

    def stencil(self):
        """Halo exchange + stencil of the band (peer memory: the exchange is hidden behind the interior rows)."""
        b = self.b
        if self.world > 1 and getattr(b, "p2p", False):
            b.stencil_p2p()
        else:
            self.exchange_input_halos()
            b.stencil()

    def seam_exchange(self):
        """Cross-band hysteresis: all-gather of the records, solve.  Returns the number of exchanges (0 or 1)."""
        b = self.b
        if self.world == 1:
            return 0
        if getattr(b, "p2p", False):
            b.seam_p2p()
        else:
            rec = b.seam_record()
            allr = b.gather_buffer(self.world)
            self.dist.all_gather_into_tensor(allr, rec, group=self.group)
            b.seam_solve(allr, self.world, self.rank)
        return 1

    def run(self):
        """Stencil + hysteresis of this band; returns the number of cross-band exchanges (0 or 1)."""
        self.stencil()
        self.b.hysteresis()
        self.exchanges = self.seam_exchange()
        return self.exchanges

    def phase_times(self, reps=5):
        """`reps` more runs with an event after every phase: median microseconds of halo exchange + stencil, band-local
        hysteresis and seam pass on this rank (waits for the peers included).  None if the backend has no events."""
        b = self.b
        if not hasattr(b, "mark"):
            return None
        names = ("stencil_us", "hysteresis_us", "seam_us")
        acc = {k: [] for k in names}
        for _ in range(reps):
            m = [b.mark()]
            for step in (self.stencil, b.hysteresis, self.seam_exchange):
                step()
                m.append(b.mark())
            b.sync()
            for i, k in enumerate(names):
                acc[k].append(1e3 * m[i].elapsed_time(m[i + 1]))
        return {k: sorted(v)[len(v) // 2] for k, v in acc.items()}


def run_local(backends, stencil=True):
    """All bands driven from ONE process (bands on the same or on different GPUs).  Same steps as BandCanny.run without
    a process group: with peer wiring (open_local) the halo rows and seam records travel through the peer-to-peer
    kernels; otherwise by tensor copies.  stencil=False: hysteresis only, on maps loaded with load_thresh.  Returns the
    number of cross-band exchanges."""
    n = len(backends)
    p2p = n > 1 and all(getattr(b, "p2p", False) for b in backends)
    for b in backends:
        b.sync()   # the uploads ran on other streams than the bands' own
    if stencil:
        if p2p:   # every push is issued before the first wait
            for b in backends:
                b.stencil_p2p(1)
            for b in backends:
                b.stencil_p2p(2)
        else:
            for i in range(n - 1):
                up, dn = backends[i], backends[i + 1]
                dn.input_rows(0, HALO).copy_(up.input_rows(up.rows, up.rows + HALO), non_blocking=True)
                up.input_rows(up.rows + HALO, up.rows + 2 * HALO).copy_(dn.input_rows(HALO, 2 * HALO), non_blocking=True)
            for b in backends:
                b.sync()
            for b in backends:
                b.stencil()
    for b in backends:
        b.hysteresis()   # (peer wiring: includes the push of the seam record)
    if n > 1:
        if p2p:
            for b in backends:
                b.seam_p2p()
        else:
            recs = [b.seam_record() for b in backends]
            for b in backends:
                b.sync()
            for r, b in enumerate(backends):
                allr = b.gather_buffer(n)
                for k, rec in enumerate(recs):
                    allr[k * b.seam_bytes:(k + 1) * b.seam_bytes].copy_(rec)
                b.seam_solve(allr, n, r)
    for b in backends:
        b.sync()
    return 1 if n > 1 else 0
