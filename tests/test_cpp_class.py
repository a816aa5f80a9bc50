"""The C++ host classes of include/b200canny.hpp (the reference's cvp::cvPipeline / cvp::cuda::CannyEdge surface over
the C ABI): compile a C++ caller with g++ (no CUDA headers), run it, compare its outputs with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
from cudacam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "canny_class_demo.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "canny_class_demo")


def _build(src=SRC, exe=EXE):
    lib_dir = os.path.join(ROOT, "cudacam_b200")
    deps = [src, os.path.join(ROOT, "include", "b200canny.hpp"), os.path.join(ROOT, "include", "b200canny.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                               "-L", lib_dir, "-lb200canny", f"-Wl,-rpath,{lib_dir}"])
    return exe


BAND_SRC = os.path.join(ROOT, "tests", "cpp", "band_runner_demo.cpp")
BAND_EXE = os.path.join(ROOT, "tests", "cpp", "band_runner_demo")


def test_cpp_caller_compiles_without_cuda_headers_and_fails_loudly_without_gpu(have_gpu, tmp_path):
    exe = _build()
    r = subprocess.run([exe, "0", "5", "64", "48", "10", "40", str(tmp_path / "o")], capture_output=True, text=True)
    if have_gpu:
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3 and "b2c::Error" in r.stderr   # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed,w,h,lo,hi", [(0, 5, 640, 360, 10, 40), (2, 9, 333, 222, 17, 43)])
def test_cpp_class_surface_matches_oracle(kind, seed, w, h, lo, hi, tmp_path):
    exe = _build()
    pre = str(tmp_path / "o")
    r = subprocess.run([exe, str(kind), str(seed), str(w), str(h), str(lo), str(hi), pre], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    f = synth.frame(kind, seed, w, h)
    want = O.canny(f, lo, hi)
    assert np.array_equal(np.fromfile(pre + ".edges", np.uint8).reshape(h, w), want["edges"])
    assert np.array_equal(np.fromfile(pre + ".blur", np.uint8).reshape(h, w), want["blur"])
    assert np.array_equal(np.fromfile(pre + ".nms", np.uint8).reshape(h, w), want["nms"])
    assert np.array_equal(np.fromfile(pre + ".grad", np.uint32).reshape(h, w), want["grad"].view(np.uint32))
    assert np.array_equal(np.fromfile(pre + ".gview", np.uint8).reshape(h, w), O.float2uchar(want["grad"]))


@pytest.mark.gpu
def test_timer_sink_keyed_by_stage_names(tmp_path):
    """the TimerManager table the UI reads (imguiApp.cpp:357-376).  Names and order are the
    reference's CANNY_STAGES (define.hpp:27-34); two HYSTER runs + one GRADIENT run after the reset -> counts 3,3,3,2,2,2;
    the fused stencil is booked on stage 1, stages 2-5 carry 0 ms, and the phases add up to the total."""
    exe = _build()
    r = subprocess.run([exe, "0", "5", "640", "360", "10", "40", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    rows = [l.split(" | ") for l in r.stdout.splitlines() if l.startswith("timer ")]
    names = [x[0][len("timer "):] for x in rows]
    assert names == ["1/6 Mono Conversion", "2/6 Gaussian Noise Removal", "3/6 Gradient Computation", "4/6 Non Maximum Suppression",
                     "5/6 Double Threshold", "6/6 Hysteresis"]
    counts = [int(x[1]) for x in rows]
    avg = [float(x[2]) for x in rows]
    assert counts == [3, 3, 3, 2, 2, 2]
    assert avg[0] > 0 and avg[5] > 0 and avg[1:5] == [0.0] * 4
    t = dict(zip(*[iter(next(l for l in r.stdout.splitlines() if l.startswith("timings ")).split()[1:])] * 2))
    t = {k: float(v) for k, v in t.items()}
    # the last run stopped at GRADIENT: no hysteresis phase; the phases tile the total (same event chain)
    assert t["upload"] > 0 and t["stencil"] > 0 and t["hysteresis"] < 0.02 and t["output"] >= 0
    assert abs(t["upload"] + t["stencil"] + t["hysteresis"] + t["output"] - t["total"]) <= 0.02 * t["total"] + 0.005
    assert abs(float(next(l for l in r.stdout.splitlines() if l.startswith("avg_hyster")).split()[1]) - avg[5]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed,w,h,nb", [(0, 3, 640, 300, 3), (1, 4, 328, 64, 2), (0, 5, 1920, 1000, 5)])
def test_cpp_band_runner(kind, seed, w, h, nb, tmp_path):
    """b2c::BandRunner (C++ driver of the row-band ABI): collective and peer-memory transports against the unsharded
    run (checked inside the demo) and against the oracle (here)."""
    exe = _build(BAND_SRC, BAND_EXE)
    pre = str(tmp_path / "b")
    r = subprocess.run([exe, str(kind), str(seed), str(w), str(h), str(nb), pre], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    want = O.canny(synth.frame(kind, seed, w, h))["edges"]
    assert np.array_equal(np.fromfile(pre + ".edges", np.uint8).reshape(h, w), want)


def test_cpp_band_runner_compiles_and_fails_loudly_without_gpu(have_gpu, tmp_path):
    exe = _build(BAND_SRC, BAND_EXE)
    if not have_gpu:
        r = subprocess.run([exe, "0", "1", "64", "64", "2", str(tmp_path / "b")], capture_output=True, text=True)
        assert r.returncode == 3 and "b2c::Error" in r.stderr
