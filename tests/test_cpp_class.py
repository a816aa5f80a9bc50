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


def _build():
    lib_dir = os.path.join(ROOT, "cudacam_b200")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "include", "b200canny.hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", EXE, SRC,
                               "-L", lib_dir, "-lb200canny", f"-Wl,-rpath,{lib_dir}"])
    return EXE


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
