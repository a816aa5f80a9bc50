"""Phase times of the union-find hysteresis (tile / border / resolve+expand kernels), measured with CUDA events between
the launches, back to back after warm-up (ncu's per-launch times are cold-cache and inflate these latency-bound kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cudacam_b200 as cb
from cudacam_b200 import _lib
lib = _lib.lib
for (w, h, n) in [(1920, 1080, 64), (1920, 1080, 8), (1920, 1080, 1)]:
    host = cb.synth.batch("scene", n, w, h, distinct=min(n, 16))
    d_in = torch.from_numpy(host.reshape(-1)).cuda()
    c = cb.CannyEdge(w, h, max_batch=max(n, 2))
    c.set_option("hyst_phase_timing", 1)
    for spread in (8, 4, 2, 1):
        c.set_option("uf_spread", spread)
        acc = np.zeros(3)
        for it in range(8):
            _lib.check(lib.b2c_run_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None, 0, 0, None))
            c.sync()
            if it >= 3:
                acc += [c.info("hyst_phase_us%d" % k) for k in range(3)]
        acc /= 5
        print(w, h, n, "spread", spread, "phase us: tile %.1f border %.1f resolve+expand %.1f  total %.1f" % (*acc, acc.sum()))
    c.close()
