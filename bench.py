#!/usr/bin/env python
"""bench.py -- Canny hot path on B200 (BASELINE.json metric: Mpixel/s, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload batch1080p|frame4k|frame720p]

One "step" = one pass of the hot path (fused stencil + on-device hysteresis -> u8 edge maps) over one batch of
synthetic frames.  Default workload = BASELINE configs[1]: 64 x 1920x1080 BGR8 frames resident in HBM (398 MB of
input per step, larger than the 126 MB L2, so no L2 flush is needed between steps).  At N GPUs every rank owns its
own batch on its own GPU (frame-parallel, no collective on the data path: "weak" scaling); the timed region is
bracketed by a barrier + synchronize, timed with CUDA events on the launching stream, max over ranks.

Keys beyond the base contract: `roofline` (fused stencil kernel, algorithmic 3.25 B/pixel, timed live with events
inside the timed region), `cpu_baseline` (OpenCV cv2 Canny chain on this box's host cores, rank 0, N=1), `e2e`
(same metric through b2c_run_batch_host with pinned HOST buffers: H2D of the frames and D2H of the edge maps inside
the timed region), `gpu_launches`, `clocks`.

`--impl reference` runs the reference's own implementation of the path: its unmodified CUDA kernels
(oracle/_ref/libcvpref.so, built from /root/reference/src/cvp/cannyEdgeD.cu) driven the way its host class drives
them (blocking upload, one launch per stage, host-driven hysteresis relaunch loop), frame by frame from host memory.
The reference has no CPU implementation of Canny.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STENCIL_BYTES_PER_PX = 3.25   # 3 B BGR8 read + 2 bits written
WORKLOADS = {
    "batch1080p": dict(w=1920, h=1080, n=64, desc="64 x 1920x1080 BGR8 synthetic 'scene' frames per GPU, device-resident (BASELINE configs[1])"),
    "frame4k": dict(w=3840, h=2160, n=1, desc="one 3840x2160 BGR8 synthetic 'scene' frame, device-resident (BASELINE configs[2])"),
    "frame720p": dict(w=1280, h=720, n=1, desc="one 1280x720 BGR8 synthetic 'scene' frame (BASELINE configs[0])"),
    # strong-scaling workloads (total work fixed as N grows); not the default bench line
    "streams1080p": dict(w=1920, h=1080, n=64, chunks_total=32, desc="8 independent 1920x1080 streams x 256 frames = 32 chunks of 64 device-resident frames, "
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

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = self.samples
        return dict(sm_mhz=(statistics.median(s) if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


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
    return dict(value=done * w * h / dt / 1e6, unit="Mpixel/s", cores=workers, kind="port",
                sample=f"OpenCV {cv2.__version__} cvtColor+GaussianBlur(5x5,1.4)+Canny(L2) on {done} frames of {w}x{h}, {workers} threads x 1 cv thread, {dt:.1f} s "
                       "(the reference has no CPU Canny; cv::Canny is the CPU baseline BASELINE.json names)")


def oracle_port_rate(frame):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    t0 = time.perf_counter()
    O.canny(frame)
    dt = time.perf_counter() - t0
    return frame.shape[0] * frame.shape[1] / dt / 1e6


def run_reference(args, wl):
    """The reference's own kernels + host-driven hysteresis loop (oracle/_ref), frame by frame from host memory."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cudacam_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    w, h, n = wl["w"], wl["h"], wl["n"]
    if not os.path.exists(O.REF_SO):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcvpref.so was not built (reference tree not mounted at build time)"}))
        return
    frames = synth.batch("scene", n, w, h, distinct=min(n, 16))
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
    peak, which = peaks()
    line = {
        "impl": "reference", "metric": "Canny edge-map throughput", "value": mpx, "unit": "Mpixel/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/fp32", "data": "synthetic",
        "config": {"workload": wl["desc"] + "; reference pipeline = unmodified src/cvp kernels, 8+k launches per frame, host-driven hysteresis loop, pageable blocking upload",
                   "hysteresis_launches_mean": float(np.mean(iters)) + 1},
        "cpu_baseline": {"value": mpx, "unit": "Mpixel/s", "cores": 1, "kind": "reference",
                         "sample": f"{args.steps} x {n} frames of {w}x{h} through oracle/_ref (the reference's CUDA kernels on this GPU, 1 host thread; it has no CPU path)"},
        # contract of the reference arm: its e2e repeats the line's own value with zero transfer bytes (the upload of
        # every frame from pageable host memory is inside the reference's run() and therefore inside `value`)
        "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "clocks": clocks,
    }
    print(json.dumps(line))


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import cudacam_b200 as cb
    from cudacam_b200 import _lib, synth
    lib = _lib.lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the Canny path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, n = wl["w"], wl["h"], wl["n"]
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
    # a real (non-default) torch stream: its handle is what the C ABI launches on, and torch.cuda.Event records on it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    assert st != 0
    H = c._h

    def stencil():
        _lib.check(lib.b2c_stencil_device(H, d_in.data_ptr(), row_stride, frame_stride, n, st), H, "stencil")

    def hyst():
        _lib.check(lib.b2c_hysteresis_device(H, n, d_edges.data_ptr(), w, w * h, st), H, "hysteresis")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        stencil()
        hyst()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    l0 = c.launches
    sampler.start()
    for a, b, e in ev:
        a.record()
        stencil()
        b.record()
        hyst()
        e.record()
        for _ in range(reps - 1):
            stencil()
            hyst()
    ev_end = torch.cuda.Event(enable_timing=True)
    ev_end.record()
    barrier()
    clocks = sampler.stop()
    launches = c.launches - l0
    total_ms = ev[0][0].elapsed_time(ev_end)
    stencil_ms = [a.elapsed_time(b) for a, b, _ in ev]
    hyst_ms = [b.elapsed_time(e) for _, b, e in ev]
    t = torch.tensor([total_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * px_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # correctness guard inside the bench: the timed output equals the host-fed public API output
    # ---- e2e: pinned host frames -> b2c_run_batch_host -> host edge maps (H2D + D2H inside the timed region) ----
    pin_in, pin_out = _lib._vp(), _lib._vp()
    _lib.check(lib.b2c_host_alloc(host.nbytes, pin_in))
    _lib.check(lib.b2c_host_alloc(px_per_chunk, pin_out))
    import ctypes as C
    C.memmove(pin_in.value, host.ctypes.data, host.nbytes)
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        for _ in range(reps):
            _lib.check(lib.b2c_run_batch_host(H, pin_in.value, row_stride, n, pin_out.value, 0), H, "run_batch_host")

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * px_per_step * e2e_steps / float(t.item()) / 1e6
    out_host = np.ctypeslib.as_array(C.cast(pin_out.value, C.POINTER(C.c_uint8)), shape=(n, h, w))
    same = bool(np.array_equal(out_host, d_edges.cpu().numpy().reshape(n, h, w)))
    edge_frac = float((out_host == 255).mean())

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
            "metric": "Canny edge-map throughput (fused stencil + on-device hysteresis)", "value": value, "unit": "Mpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if "chunks_total" in wl else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["desc"], "frames_per_gpu": n, "width": w, "height": h, "thresholds": [10, 40],
                       "l2_policy": "inputs (%.0f MB/launch/GPU) larger than the 126 MB L2, no flush" % (px_per_chunk * 3 / 1e6) if px_per_chunk * 3 > 130e6 else "input smaller than L2: latency workload, L2-warm",
                       "parallelism": f"frame-parallel x{world}, no collective", "stencil_impl": c.info("stencil_impl"),
                       "e2e_equals_device_path": same, "edge_pixel_fraction": edge_frac},
            "roofline": {"bound": "hbm", "kernel": "fused stencil (BGR8 -> 2-bit weak/strong map)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which, "algorithmic_bytes_per_launch": px_per_chunk * STENCIL_BYTES_PER_PX,
                         "kernel_ms": k_ms, "hysteresis_ms": statistics.mean(hyst_ms), "stencil_share_of_step": sum(stencil_ms) / (sum(stencil_ms) + sum(hyst_ms))},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": px_per_step * 3, "d2h_bytes_per_step": px_per_step, "steps": e2e_steps,
                    "api": "b2c_run_batch_host (pinned host frames in, host u8 edge maps out)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_cv2([host[i] for i in range(distinct)])
            line["cpu_baseline"]["oracle_port_mpixel_s_1core"] = oracle_port_rate(host[0])
        if args.workload == "frame4k":
            lat = sorted(a.elapsed_time(e) for a, _, e in ev)
            line["latency_ms"] = {"median": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))], "min": lat[0]}
        print(json.dumps(line))
    lib.b2c_host_free(pin_in)
    lib.b2c_host_free(pin_out)
    c.close()
    if world > 1:
        dist.destroy_process_group()


def run_giga(args, wl):
    """BASELINE configs[4]: one 16384x16384 image, one row band per rank (strong scaling).  A step = 4-row input halo
    exchange with the neighbour ranks (peer stores over NVLink, or NCCL send/recv), fused stencil on the band, then band-local hysteresis +
    boundary-row exchange + convergence flags (device side) or all-reduce (NCCL) until the global fixpoint; every
    resolve pass keeps the u8 edge map current."""
    import torch
    import torch.distributed as dist
    from cudacam_b200 import bands, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the Canny path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, Hh = wl["w"], wl["h"]
    y0, rows = bands.band_rows(Hh, world, rank)
    band_host = synth.giga_rows(y0, y0 + rows, W, Hh)
    be = bands.CudaBandBackend(W, rows, y0, Hh, device=local)
    be.load(band_host)
    p2p = world > 1 and os.environ.get("B2C_BAND_NCCL", "0") != "1"
    if p2p:
        be.enable_p2p(dist, rank, world)   # cross-band rounds on the devices (NVLink peer stores); B2C_BAND_NCCL=1 = NCCL rounds
    bc = bands.BandCanny(be, rank, world, dist if world > 1 else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        bc.run()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = be.launches
    a.record()
    rounds = 0
    for _ in range(args.steps):
        rounds = bc.run()
    b.record()
    barrier()
    clocks = sampler.stop()
    launches = be.launches - l0
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # e2e: band from host memory, edge map back to host, every step
    e2e_steps = min(args.steps, 3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        be.load(band_host)
        bc.run()
        edges = be.edges()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, which = peaks()
        px = W * Hh
        line = {
            "metric": "Canny edge-map throughput (row-band sharded gigapixel image)", "value": px * args.steps / (total_ms * 1e-3) / 1e6, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["desc"], "width": W, "height": Hh, "band_rows_rank0": rows, "thresholds": [10, 40], "global_hysteresis_rounds": rounds,
                       "l2_policy": "band input (%.0f MB) larger than the 126 MB L2, no flush" % (rows * W * 3 / 1e6), "parallelism": f"row bands x{world}, NCCL input halo; hysteresis rounds: " + ("peer-memory stores + device-side convergence flags" if p2p else "NCCL send/recv + 1-int all-reduce per round"),
                       "edge_pixel_fraction_rank0": float((edges == 255).mean())},
            "roofline": {"bound": "hbm", "kernel": "whole step (4 B/pixel end to end: 3 in + 1 out)", "achieved": px * 4.0 / (total_ms / args.steps * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": px * 4.0 / (total_ms / args.steps * 1e-3) / 1e9 / world / peak, "traffic": None, "peak_source": which},
            "e2e": {"value": px * e2e_steps / float(t.item()) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": px * 3, "d2h_bytes_per_step": px, "steps": e2e_steps,
                    "api": "CudaBandBackend.load (pageable host band) + BandCanny.run + edges() download"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    be.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="batch1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if args.steps > 10:
            args.steps = 10   # bounded: the reference needs ~1 ms per frame
        run_reference(args, wl)
    elif args.workload == "giga":
        run_giga(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
