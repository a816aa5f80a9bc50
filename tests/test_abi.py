"""The C-ABI library loads and exports every symbol include/b200canny.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200canny.h")).read()
    return sorted(set(re.findall(r"B2C_API\s+[\w\s\*]+?\b(b2c_\w+)\s*\(", src)))


def test_header_symbols_exported():
    from cudacam_b200 import _lib
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/b200canny.h but not exported"
    # and the ctypes table covers the header exactly
    assert sorted(_lib.EXPORTS) == names


def test_no_oracle_in_product():
    """The product library must not reference the oracle, and the package must not import it."""
    from cudacam_b200 import _lib
    import subprocess
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "cvpref_" not in out
    for f in os.listdir(os.path.join(ROOT, "cudacam_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "cudacam_b200", f)).read()


def test_status_strings_and_version():
    from cudacam_b200 import _lib
    assert _lib.lib.b2c_strerror(0) == b"ok"
    assert b"geometry" in _lib.lib.b2c_strerror(_lib.ERR_SIZE)
    assert b"sm_100a" in _lib.lib.b2c_version()


def test_argument_errors_without_gpu():
    from cudacam_b200 import _lib
    h = C.c_void_p()
    assert _lib.lib.b2c_create(C.byref(h), 0, 0, 10, 3, 1) == _lib.ERR_INVALID
    assert _lib.lib.b2c_create(C.byref(h), 0, 64, 48, 2, 1) == _lib.ERR_UNSUPPORTED   # channels must be 1 (GRAY8), 3 (BGR8) or 4 (BGRA8)
    assert _lib.lib.b2c_create(C.byref(h), 0, 64, 48, 0x104, 1) == _lib.ERR_UNSUPPORTED   # only BGR8 has a planar form (0x103)
    assert _lib.lib.b2c_create(None, 0, 64, 48, 3, 1) == _lib.ERR_INVALID
    assert _lib.lib.b2c_run(None, None, 0, 5) == _lib.ERR_INVALID
    assert _lib.lib.b2c_get_low_threshold(None) == _lib.ERR_INVALID


def test_create_fails_loudly_without_device(have_gpu):
    if have_gpu:
        pytest.skip("device present")
    import cudacam_b200 as cb
    with pytest.raises(cb.B2cError):
        cb.CannyEdge(64, 48)


def test_synth_deterministic():
    from cudacam_b200 import synth
    a = synth.frame("scene", 5, 160, 90)
    b = synth.frame("scene", 5, 160, 90)
    c = synth.frame("scene", 6, 160, 90)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert synth.frame("noise", 1, 33, 17).shape == (17, 33, 3)
    s = synth.frame("steps", 1, 200, 100)
    assert (s == 0).any() and (s == 255).any()
    # strided output rows
    buf = np.zeros((90, 512), np.uint8)
    from cudacam_b200 import _lib
    assert _lib.lib.b2c_synth_frame(0, 5, 160, 90, buf.ctypes.data, 512) == 0
    assert np.array_equal(buf[:, :480].reshape(90, 160, 3), a)
