"""Times the marching stencil kernel for several rows-per-band values and start staggers (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cudacam_b200 as cb
from cudacam_b200 import _lib
lib = _lib.lib
cfgs = [(1920, 1080, 64, [36, 96, 136, 276, 546], [0, 1000, 2000, 4000, 8000, 16000]), (3840, 2160, 1, [16, 26], [0, 1000, 4000])]
for (w, h, n, rbs, sts) in cfgs:
    host = cb.synth.batch("scene", n, w, h, distinct=min(n, 8))
    d_in = torch.from_numpy(host.reshape(-1)).cuda()
    c = cb.CannyEdge(w, h, max_batch=max(n, 2))
    for rb in rbs:
        for st in sts:
            c.set_option("march_rb", rb)
            c.set_option("march_stagger_ns", st)
            for it in range(3):
                _lib.check(lib.b2c_stencil_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None))
            c.sync()
            t0 = time.perf_counter()
            for it in range(20):
                _lib.check(lib.b2c_stencil_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None))
            c.sync()
            dt = (time.perf_counter() - t0) / 20
            print(w, h, n, "rb", rb, "stagger", st, "us %.1f" % (dt * 1e6), flush=True)
    c.close()
