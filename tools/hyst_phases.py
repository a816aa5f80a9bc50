import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cudacam_b200 as cb
from cudacam_b200 import _lib
lib = _lib.lib
for (w, h, n) in [(3840, 2160, 1), (1920, 1080, 64), (1920, 1080, 1)]:
    host = cb.synth.batch("scene", n, w, h, distinct=min(n, 8))
    d_in = torch.from_numpy(host.reshape(-1)).cuda()
    c = cb.CannyEdge(w, h, max_batch=max(n, 2))
    for it in range(3):
        _lib.check(lib.b2c_run_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None, 0, 0, None))
        c.sync()
    st = [c.info("stamp%d" % k) & 0xFFFFFFFF for k in range(6)]
    d = [((st[k + 1] - st[k]) & 0xFFFFFFFF) / 1000.0 for k in range(5)]
    print(w, h, n, "phase us: planes %.1f init %.1f union %.1f resolve %.1f expand %.1f  total %.1f" % (*d, sum(d)), "grid", c.info("hyst_grid"))
    c.close()
