#!/usr/bin/env python
"""bench.py -- Canny hot path on B200 (BASELINE.json metric: Mpixel/s, 4K frame latency, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-cpu] [--no-extras]

One "step" = one pass of the hot path (fused stencil + on-device hysteresis -> u8 edge maps) over one batch of
synthetic frames.  Default workload = BASELINE configs[1]: 64 x 1920x1080 BGR8 frames per GPU (398 MB of input per
step, larger than the 126 MB L2, so no L2 flush is needed between steps).  At N GPUs every rank owns its own batch on
its own GPU (frame-parallel, no collective on the data path: "weak" scaling); the timed region is bracketed by a
barrier + synchronize, timed with CUDA events on the launching stream, max over ranks.

Both arms print the SAME `metric`, `unit` and `config.workload`; what differs is `impl` and `config.residency`.

Keys beyond the base contract:
  roofline      fused stencil kernel, algorithmic 3.25 B/pixel, timed live with events inside the timed region
  e2e           same metric through b2c_run_batch_host with pinned HOST buffers (H2D of the frames and D2H of the edge
                maps inside the timed region), plus the packed 1-bit variant and the box's measured PCIe ceilings
  oracle_check  frame 0 of the timed batch against the CPU oracle (bit-exact or the run fails)
  cpu_baseline  OpenCV cv2 Canny chain on this box's host cores (rank 0, N = 1)
  latency_4k    BASELINE configs[2] (N = 1): one 3840x2160 frame device-resident in -> u8 edge map out, on-device
                hysteresis; stream launches, CUDA-graph replay (the capture itself proves there is no host round trip
                on the path) and a cold-L2 variant
  giga          BASELINE configs[4] (every N): one 16384x16384 image in N row bands, halo exchange + cross-band
                hysteresis to the global fixpoint; sha256 of the assembled edge map (must be the same at every N and
                equal to the oracle's, tests/golden/giga_sha256.json)

`--impl reference` runs the reference's own implementation of the path: its unmodified CUDA kernels
(oracle/_ref/libcvpref.so, built from /root/reference/src/cvp/cannyEdgeD.cu) driven the way its host class drives
them (blocking pageable upload, one launch per stage, host-driven hysteresis relaunch loop), frame by frame from host
memory, on the same workload and with the same --steps.  The reference has no CPU implementation of Canny and no
multi-GPU mode: under torchrun rank 0 alone runs it.  The product library is NOT loaded in that process (frames come
from the stand-alone generator library libb200synth.so).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Canny edge-map throughput"
UNIT = "Mpixel/s"
STENCIL_BYTES_PER_PX = 3.25   # 3 B BGR8 read + 2 bits written
WORKLOADS = {
    "batch1080p": dict(w=1920, h=1080, n=64, desc="64 x 1920x1080 BGR8 synthetic 'scene' frames per GPU (BASELINE configs[1])"),
    "frame4k": dict(w=3840, h=2160, n=1, desc="one 3840x2160 BGR8 synthetic 'scene' frame (BASELINE configs[2])"),
    "frame720p": dict(w=1280, h=720, n=1, desc="one 1280x720 BGR8 synthetic 'scene' frame (BASELINE configs[0])"),
    # strong-scaling workloads (total work fixed as N grows); not the default bench line
    "streams1080p": dict(w=1920, h=1080, n=64, chunks_total=32, desc="8 independent 1920x1080 streams x 256 frames = 32 chunks of 64 frames, "
                         "stream s on GPU s mod N, no collective (BASELINE configs[3])"),
    "giga": dict(w=16384, h=16384, n=1, desc="one 16384x16384 BGR8 synthetic mosaic, row bands over the ranks, 4-row input halo exchange + cross-band hysteresis "
                 "to a global fixpoint (BASELINE configs[4])"),
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def sample(self):
        """One reading.  The timed loops call it themselves once all steps are issued and the GPU is still working through
        them (a reading inside the timed region at no cost to it); the thread adds more on long runs."""
        nv = self.nv
        if not nv:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
            for bit, name in self.NAMES.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            time.sleep(0.02)   # (NVML queries take a driver lock: polling harder than this slows the launches of all ranks)
            if not self._stop.is_set():
                self.sample()

    def start(self):
        if self.nv and os.environ.get("B2C_NO_SAMPLER", "0") != "1":
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = self.samples
        return dict(sm_mhz=(statistics.median(s) if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


# ---- CPU baseline -------------------------------------------------------------------------------------------------
def cpu_baseline_cv2(frames, lo=10, hi=40, budget_s=12.0):
    """OpenCV chain named by BASELINE.json (gray -> 5x5 Gaussian sigma 1.4 -> Canny L2, thresholds mapped as in
    BASELINE.md 2a), frame-parallel over all host cores with cv threads = 1 each (cv2 releases the GIL)."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    cv2.setNumThreads(1)
    tl = float(np.sqrt((2 * lo + 2) ** 2 - 0.5))
    th = float(np.sqrt((2 * hi + 2) ** 2 - 0.5))

    def one(i):
        g = cv2.cvtColor(frames[i % len(frames)], cv2.COLOR_BGR2GRAY)
        b = cv2.GaussianBlur(g, (5, 5), 1.4)
        return int(cv2.Canny(b, tl, th, apertureSize=3, L2gradient=True)[0, 0])

    h, w = frames[0].shape[:2]
    workers = min(cores, 64)
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(one, range(workers)))   # warm-up
        done, t0 = 0, time.perf_counter()
        while True:
            list(ex.map(one, range(workers * 2)))
            done += workers * 2
            dt = time.perf_counter() - t0
            if dt > budget_s or done >= 4096:
                break
    return dict(value=done * w * h / dt / 1e6, unit=UNIT, cores=workers, kind="port",
                sample=f"OpenCV {cv2.__version__} cvtColor+GaussianBlur(5x5,1.4)+Canny(L2) on {done} frames of {w}x{h}, {workers} threads x 1 cv thread, {dt:.1f} s "
                       "(the reference has no CPU Canny; cv::Canny is the CPU baseline BASELINE.json names)")


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    return O


# ---- frames without the product library (reference arm) -------------------------------------------------------------
def synth_standalone_batch(n, w, h, distinct, stream=0):
    """Same frames as cudacam_b200.synth.batch("scene", ...), from libb200synth.so (csrc/synth.cpp alone, no CUDA)."""
    lib = C.CDLL(os.path.join(ROOT, "cudacam_b200", "libb200synth.so"))
    lib.b2c_synth_frame.restype = C.c_int
    lib.b2c_synth_frame.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    out = np.empty((n, h, w, 3), np.uint8)
    for f in range(min(distinct, n)):
        seed = (0xC0FFEE ^ (stream << 32) ^ f) & (2**64 - 1)
        if lib.b2c_synth_frame(0, seed, w, h, out[f].ctypes.data, w * 3) != 0:
            raise SystemExit("b2c_synth_frame failed")
    for f in range(distinct, n):
        out[f] = out[f % distinct]
    return out


def run_reference(args, wl):
    """The reference's own kernels + host-driven hysteresis loop (oracle/_ref), frame by frame from host memory."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O = _oracle()
    w, h, n = wl["w"], wl["h"], wl["n"]
    if args.workload in ("giga", "streams1080p"):
        print(json.dumps({"impl": "reference", "unavailable": f"the reference has no multi-GPU / batched mode for workload {args.workload}"}))
        return
    if not os.path.exists(O.REF_SO):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcvpref.so was not built (reference tree not mounted at build time)"}))
        return
    frames = synth_standalone_batch(n, w, h, distinct=min(n, 16))
    ref = O.CvpRef(w, h)
    sampler = ClockSampler(0)
    iters = []

    def step():
        for i in range(n):
            ref.run(frames[i], 5)
            iters.append(ref.info()[0])

    for _ in range(args.warmup):
        step()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0   # cvpref_run ends with cudaDeviceSynchronize: wall clock == device-complete
    clocks = sampler.stop()
    ref.close()
    mpx = args.steps * n * w * h / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mpx, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl["desc"], "frames_per_gpu": n, "width": w, "height": h, "thresholds": [10, 40],
                   "residency": "host-fed: every frame is uploaded from pageable host memory by the reference's own run() (blocking cudaMemcpy2D), inside the timed region",
                   "pipeline": "unmodified src/cvp kernels (oracle/_ref), 8+k launches per frame, host-driven hysteresis relaunch loop with 2 blocking 4-byte copies per round",
                   "hysteresis_launches_mean": float(np.mean(iters)) + 1, "arithmetic": "u8 / fp32, as the reference"},
        "cpu_baseline": {"value": mpx, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": f"{args.steps} x {n} frames of {w}x{h} through oracle/_ref (the reference's CUDA kernels on this GPU, 1 host thread; it has no CPU path)"},
        # contract of the reference arm: its e2e repeats the line's own value with zero transfer bytes declared (the
        # upload of every frame is inside the reference's run() and therefore already inside `value`)
        "e2e": {"value": mpx, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "clocks": clocks,
    }
    print(json.dumps(line))


# ---- helpers of the product arm -----------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the Canny path")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        # a real (non-default) torch stream: its handle is what the C ABI launches on, and torch.cuda.Event records on it
        self.tstream = torch.cuda.Stream()
        torch.cuda.set_stream(self.tstream)
        self.st = self.tstream.cuda_stream
        assert self.st != 0

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def pct(sorted_vals, q):
    return sorted_vals[min(len(sorted_vals) - 1, int(len(sorted_vals) * q))]


def pcie_ceiling(env, nbytes_in, nbytes_out, reps=6):
    """The box's own host<->device copy rates with bare cudaMemcpyAsync calls on pinned memory (one call per copy):
    H2D alone, D2H alone, and both at once on two streams -- what the e2e number has to be read against."""
    torch = env.torch
    hin = torch.empty(nbytes_in, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes_out, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes_in, dtype=torch.uint8, device="cuda")
    dout = torch.empty(nbytes_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)

    def both():
        h2d()
        d2h()

    for f in (h2d, d2h):
        f()
    t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
    return {"h2d_gbs": nbytes_in / t_in / 1e9, "d2h_gbs": nbytes_out / t_out / 1e9,
            "concurrent_h2d_gbs": nbytes_in / t_both / 1e9, "concurrent_d2h_gbs": nbytes_out / t_both / 1e9,
            "how": f"cudaMemcpyAsync on pinned memory, {nbytes_in >> 20} MiB up / {nbytes_out >> 20} MiB down per call, mean of {reps}"}


def measure_latency_4k(env, runs=200):
    """BASELINE configs[2]: one 3840x2160 frame, device-resident in -> final u8 edge map on the device."""
    import cudacam_b200 as cb
    from cudacam_b200 import _lib, synth
    torch, lib = env.torch, _lib.lib
    w, h = 3840, 2160
    nf = 6   # 6 x 24.9 MB = 149 MB of distinct inputs > 126 MB L2
    host = synth.batch("scene", nf, w, h, stream=77, distinct=nf)
    d_in = torch.from_numpy(host.reshape(nf, -1)).cuda()
    d_edges = torch.empty(h * w, dtype=torch.uint8, device="cuda")
    c = cb.CannyEdge(w, h, device=env.local, max_batch=1)
    c.enableKernelProfiling(False)
    H = c._h

    def run(i):
        _lib.check(lib.b2c_run_device(H, d_in[i % nf].data_ptr(), w * 3, w * 3 * h, 1, d_edges.data_ptr(), w, w * h, env.st), H, "run_device")

    def series(fn, n, before=None):
        ev = []
        for i in range(n):
            if before:
                before()
            a, b = env.event(), env.event()
            a.record()
            fn(i)
            b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        v = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
        return {"median_us": pct(v, 0.5), "p99_us": pct(v, 0.99), "min_us": v[0], "runs": n}

    for i in range(10):
        run(i)
    torch.cuda.synchronize()
    l0 = c.launches
    run(0)
    launches_per_frame = int(c.launches - l0)
    out = {"width": w, "height": h, "launches_per_frame": launches_per_frame,
           "same_frame_l2_warm": series(lambda i: run(0), runs),
           "rotating_frames": series(run, runs)}
    # CUDA graph of one frame: stream capture refuses any synchronising call, so a successful capture is the proof
    # that nothing on the path goes back to the host; replay also removes the launch gaps
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=env.tstream):
            run(0)
        for _ in range(5):
            g.replay()
        out["graph_replay"] = series(lambda i: g.replay(), runs)
        out["graph_replay"]["captured"] = True
        out["host_round_trips_on_path"] = 0
    except Exception as e:   # report, do not hide
        out["graph_replay"] = {"captured": False, "error": str(e)[:200]}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out["cold_l2"] = series(run, 40, before=lambda: flush.fill_(1))
    out["cold_l2"]["how"] = "512 MiB fill before every timed frame (inputs, map, planes and edge map all evicted)"
    # parity of the timed output
    O = _oracle()
    run(1)
    torch.cuda.synchronize()
    got = d_edges.cpu().numpy().reshape(h, w)
    out["oracle_equal"] = bool(np.array_equal(got, O.canny(host[1])["edges"]))
    out["target_us"] = 200
    c.close()
    return out


def measure_giga(env, steps, warmup, W=16384, Hh=16384, want_edges=True):
    """BASELINE configs[4]: one image in `world` row bands.  Returns (record, total_ms, launches, clocks)."""
    from cudacam_b200 import bands, synth
    torch, dist = env.torch, env.dist
    world, rank = env.world, env.rank
    y0, rows = bands.band_rows(Hh, world, rank)
    band_host = synth.giga_rows(y0, y0 + rows, W, Hh)
    be = bands.CudaBandBackend(W, rows, y0, Hh, device=env.local)
    be.load(band_host)
    p2p = world > 1 and os.environ.get("B2C_BAND_NCCL", "0") != "1"
    if p2p:   # cross-band exchange on the devices (NVLink peer stores); B2C_BAND_NCCL=1, or no CUDA IPC on this box: NCCL
        p2p = be.enable_p2p(dist, rank, world)
    bc = bands.BandCanny(be, rank, world, dist if world > 1 else None)
    for _ in range(max(warmup, 3)):
        bc.run()
    sampler = ClockSampler(env.local)
    sampler.start()   # (before the barrier: the thread's start-up and first NVML queries are not part of the timed region)
    env.barrier()
    a, b = env.event(), env.event()
    l0 = be.launches
    a.record()
    rounds = 0
    for _ in range(steps):
        rounds = bc.run()
    b.record()
    sampler.sample()   # all steps are issued, the GPU is still running them
    env.barrier()
    clocks = sampler.stop()
    launches = be.launches - l0
    total_ms = env.max_over_ranks(a.elapsed_time(b))
    phases = bc.phase_times() if hasattr(bc, "phase_times") else None
    # e2e: band from host memory, edge map back to host, every step
    e2e_steps = min(steps, 3)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        be.load(band_host)
        bc.run()
        edges = be.edges()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    # the assembled edge map's hash on rank 0 (same bytes at every N, or the sharding is wrong)
    digest = None
    if want_edges:
        mine = torch.from_numpy(edges).cuda()
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
            if Hh % world:
                raise SystemExit("giga: rows must split evenly for the gather")
            dist.gather(mine, parts, dst=0)
            full = torch.cat(parts).cpu().numpy() if rank == 0 else None
        else:
            full = edges
        if rank == 0:
            digest = hashlib.sha256(np.ascontiguousarray(full).tobytes()).hexdigest()
    px = W * Hh
    rec = None
    if rank == 0:
        golden = None
        try:
            golden = json.load(open(os.path.join(ROOT, "tests", "golden", "giga_sha256.json"))).get(f"{W}x{Hh}")
        except Exception:
            pass
        rec = {"workload": WORKLOADS["giga"]["desc"], "width": W, "height": Hh, "bands": world, "band_rows_rank0": rows, "scaling": "strong",
               "ms_per_step": total_ms / steps, "value": px * steps / (total_ms * 1e-3) / 1e6, "unit": UNIT, "steps": steps,
               "global_hysteresis_exchanges": rounds, "phase_us_rank0": phases,
               "protocol": ("peer-memory stores over NVLink + device-side seam solve" if p2p else "NCCL send/recv of the halo rows + all-gather of the seam records" if world > 1 else "single band"),
               "e2e_value": px * e2e_steps / e2e_s / 1e6, "edge_pixel_fraction_rank0": float((edges == 255).mean()),
               "sha256_edges": digest, "sha256_oracle_golden": golden, "equals_oracle_golden": (digest == golden) if (digest and golden) else None,
               "hbm_frac_whole_step": px * 4.0 / (total_ms / steps * 1e-3) / 1e9 / world / peaks()[0]}
    be.close()
    return rec, total_ms, launches, clocks


def run_ours(args, wl):
    env = Env()
    torch = env.torch
    import cudacam_b200 as cb
    from cudacam_b200 import _lib, synth
    lib = _lib.lib
    world, rank, local = env.world, env.rank, env.local
    # one process per GPU: keep this rank's threads and its pinned rings on the NUMA node of its GPU (-1: platform silent)
    # (only with several ranks: the N=1 run also times the CPU baseline on ALL host cores)
    numa_node = lib.b2c_bind_host_to_device(local) if world > 1 else None
    w, h, n = wl["w"], wl["h"], wl["n"]
    warmup = max(args.warmup, 3)
    reps = 1
    if "chunks_total" in wl:
        if wl["chunks_total"] % world:
            raise SystemExit(f"{args.workload}: {wl['chunks_total']} chunks do not split over {world} ranks")
        reps = wl["chunks_total"] // world
    px_per_chunk = n * w * h
    px_per_step = px_per_chunk * reps

    # synthetic frames: stream id = rank, so every GPU works on different pictures
    distinct = min(n, 16)
    host = synth.batch("scene", n, w, h, stream=rank, distinct=distinct)
    row_stride = w * 3
    frame_stride = row_stride * h
    d_in = torch.from_numpy(host.reshape(-1)).cuda()
    d_edges = torch.empty(n * h * w, dtype=torch.uint8, device="cuda")
    c = cb.CannyEdge(w, h, device=local, max_batch=max(n, 2))
    c.enableKernelProfiling(False)
    st, H = env.st, c._h

    def stencil():
        _lib.check(lib.b2c_stencil_device(H, d_in.data_ptr(), row_stride, frame_stride, n, st), H, "stencil")

    def hyst():
        _lib.check(lib.b2c_hysteresis_device(H, n, d_edges.data_ptr(), w, w * h, st), H, "hysteresis")

    for _ in range(warmup):
        stencil()
        hyst()
    ev = [(env.event(), env.event(), env.event()) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()   # (before the barrier: the thread's start-up is not part of the timed region)
    env.barrier()
    l0 = c.launches
    for a, b, e in ev:
        a.record()
        stencil()
        b.record()
        hyst()
        e.record()
        for _ in range(reps - 1):
            stencil()
            hyst()
    ev_end = env.event()
    ev_end.record()
    sampler.sample()   # all steps are issued, the GPU is still running them
    env.barrier()
    clocks = sampler.stop()
    launches = c.launches - l0
    stencil_ms = [a.elapsed_time(b) for a, b, _ in ev]
    hyst_ms = [b.elapsed_time(e) for _, b, e in ev]
    total_ms = env.max_over_ranks(ev[0][0].elapsed_time(ev_end))
    value = world * px_per_step * args.steps / (total_ms * 1e-3) / 1e6
    dev_edges_host = d_edges.cpu().numpy().reshape(n, h, w)

    # ---- oracle check: frame 0 of the timed batch against the CPU oracle (test infrastructure used as the checker)
    oracle_check = None
    if rank == 0:
        O = _oracle()
        t0 = time.perf_counter()
        want = O.canny(host[0])["edges"]
        oracle_s = time.perf_counter() - t0
        ok = bool(np.array_equal(dev_edges_host[0], want))
        oracle_check = {"frame": 0, "equal": ok, "edge_pixels": int((want == 255).sum()), "oracle_port_mpixel_s_1core": w * h / oracle_s / 1e6}
        if not ok:
            raise SystemExit("bench: the timed batch's frame 0 differs from the oracle -- numbers withheld")

    # ---- e2e: pinned host frames -> b2c_run_batch_host -> host edge maps (H2D + D2H inside the timed region) ----
    pin_in, pin_out = _lib._vp(), _lib._vp()
    _lib.check(lib.b2c_host_alloc(host.nbytes, pin_in))
    _lib.check(lib.b2c_host_alloc(px_per_chunk, pin_out))
    C.memmove(pin_in.value, host.ctypes.data, host.nbytes)
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_run(packed):
        def step():
            for _ in range(reps):
                _lib.check(lib.b2c_run_batch_host(H, pin_in.value, row_stride, n, pin_out.value, packed), H, "run_batch_host")
        for _ in range(2):
            step()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step()
        torch.cuda.synchronize()
        return world * px_per_step * e2e_steps / env.max_over_ranks(time.perf_counter() - t0) / 1e6

    e2e_bits = e2e_run(1)
    wpr = (w + 31) // 32
    bits_host = np.ctypeslib.as_array(C.cast(pin_out.value, C.POINTER(C.c_uint32)), shape=(n, h, wpr)).copy()
    e2e_value = e2e_run(0)
    out_host = np.ctypeslib.as_array(C.cast(pin_out.value, C.POINTER(C.c_uint8)), shape=(n, h, w))
    same = bool(np.array_equal(out_host, dev_edges_host))
    same_bits = bool(np.array_equal(np.unpackbits(bits_host.view(np.uint8), axis=-1, bitorder="little")[:, :, :w] * 255, dev_edges_host))
    edge_frac = float((out_host == 255).mean())
    # every rank measures at the same time (barrier first): with N ranks on one box this is the per-GPU share of the
    # host's PCIe complex under the same load pattern as the e2e run, not the link speed of one idle GPU
    env.barrier()
    ceiling = pcie_ceiling(env, px_per_chunk * 3, px_per_chunk)
    if ceiling:
        ceiling["ranks_copying_at_once"] = world
    lib.b2c_host_free(pin_in)
    lib.b2c_host_free(pin_out)
    stencil_impl = c.info("stencil_impl")
    c.close()
    del d_in, d_edges
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        peak, which = peaks()
        k_ms = statistics.mean(stencil_ms)
        achieved = px_per_chunk * STENCIL_BYTES_PER_PX / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "stencil_traffic.json"))).get("bytes_per_launch_" + args.workload)
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if "chunks_total" in wl else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "b200canny",
            "config": {"workload": wl["desc"], "frames_per_gpu": n, "width": w, "height": h, "thresholds": [10, 40],
                       "residency": "`value`: frames resident in HBM when the timed region starts; `e2e`: pinned host frames in, host edge maps out, copies inside the timed region",
                       "pipeline": "fused stencil (1 launch) + on-device union-find hysteresis, no host round trip",
                       "l2_policy": "inputs (%.0f MB/launch/GPU) larger than the 126 MB L2, no flush" % (px_per_chunk * 3 / 1e6) if px_per_chunk * 3 > 130e6 else "input smaller than L2: latency workload, L2-warm",
                       "parallelism": f"frame-parallel x{world}, no collective", "stencil_impl": stencil_impl,
                       "e2e_equals_device_path": same, "e2e_bits_equal_device_path": same_bits, "edge_pixel_fraction": edge_frac},
            "roofline": {"bound": "hbm", "kernel": "fused stencil k_stencil_march (BGR8 -> strong and weak|strong bit planes)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which, "algorithmic_bytes_per_launch": px_per_chunk * STENCIL_BYTES_PER_PX,
                         "kernel_ms": k_ms, "hysteresis_ms": statistics.mean(hyst_ms), "stencil_share_of_step": sum(stencil_ms) / (sum(stencil_ms) + sum(hyst_ms))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": px_per_step * 3, "d2h_bytes_per_step": px_per_step, "steps": e2e_steps,
                    "api": "b2c_run_batch_host (pinned host frames in, host u8 edge maps out; 8 slots of 8 frames on 3 streams)", "numa_node_rank0": numa_node,
                    "h2d_gbs_per_gpu": e2e_value * 3 / 1e3 / world, "packed_bits_value": e2e_bits, "packed_bits_d2h_bytes_per_step": n * h * wpr * 4 * reps,
                    "pcie_ceiling_rank0": ceiling},
            "oracle_check": oracle_check,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if ceiling:
            line["e2e"]["frac_of_h2d_ceiling"] = line["e2e"]["h2d_gbs_per_gpu"] / ceiling["h2d_gbs"]
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_cv2([host[i] for i in range(distinct)])
            line["cpu_baseline"]["oracle_port_mpixel_s_1core"] = oracle_check["oracle_port_mpixel_s_1core"]
        if args.workload == "frame4k":
            lat = sorted(a.elapsed_time(e) for a, _, e in ev)
            line["latency_ms"] = {"median": pct(lat, 0.5), "p99": pct(lat, 0.99), "min": lat[0]}
    del host

    # ---- sub-records of the other BASELINE configs, so that the driver's own runs carry them -------------------------
    if args.workload == "batch1080p" and not args.no_extras:
        if world == 1:
            lat = measure_latency_4k(env)
            line["latency_4k"] = lat
        # (at least 50 steps: a step is 0.25-0.9 ms, and one scheduling hiccup of a rank stalls all ranks of the band protocol)
        rec, _, _, _ = measure_giga(env, steps=max(50, args.steps), warmup=5)
        if rank == 0:
            line["giga"] = rec
    if rank == 0:
        print(json.dumps(line))
    env.close()


def run_giga(args, wl):
    """BASELINE configs[4] as its own bench line (strong scaling over the ranks)."""
    env = Env()
    if os.environ.get("B2C_GIGA_H"):   # experiments only: a shorter image (bands as small as at 8 ranks on fewer GPUs)
        wl = dict(wl, h=int(os.environ["B2C_GIGA_H"]))
    rec, total_ms, launches, clocks = measure_giga(env, args.steps, args.warmup, wl["w"], wl["h"])
    if env.rank == 0:
        peak, which = peaks()
        px = wl["w"] * wl["h"]
        ach = px * 4.0 / (total_ms / args.steps * 1e-3) / 1e9 / env.world
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "b200canny",
            "config": {"workload": wl["desc"], "width": wl["w"], "height": wl["h"], "thresholds": [10, 40],
                       "l2_policy": "band input (%.0f MB) larger than the 126 MB L2, no flush" % (rec["band_rows_rank0"] * wl["w"] * 3 / 1e6),
                       "parallelism": f"row bands x{env.world}; " + rec["protocol"]},
            "roofline": {"bound": "hbm", "kernel": "whole step (4 B/pixel end to end: 3 in + 1 out)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": which},
            "e2e": {"value": rec["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": px * 3, "d2h_bytes_per_step": px,
                    "api": "CudaBandBackend.load (pageable host band) + BandCanny.run + edges() download"},
            "giga": rec, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the latency_4k / giga sub-records of the default line")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif args.workload == "giga":
        run_giga(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
