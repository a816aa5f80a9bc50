"""Times the marching stencil kernel for several rows-per-band values (GPU box).  usage: sweep_rb.py [rb ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cudacam_b200 as cb
from cudacam_b200 import _lib
lib = _lib.lib
rbs = [int(a) for a in sys.argv[1:]] or [20, 32, 44, 56, 68, 92, 116, 140, 176, 272]
cfgs = [(1920, 1080, 64, rbs), (3840, 2160, 1, [8, 20, 32, 44])]
for (w, h, n, rbl) in cfgs:
    host = cb.synth.batch("scene", n, w, h, distinct=min(n, 8))
    d_in = torch.from_numpy(host.reshape(-1)).cuda()
    c = cb.CannyEdge(w, h, max_batch=max(n, 2))
    print(w, h, n, "ctas/sm", c.info("march_ctas_per_sm"), "auto rb", c.info("march_band_rows"), flush=True)
    for rb in rbl:
        c.set_option("march_rb", rb)
        for it in range(3):
            _lib.check(lib.b2c_stencil_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None))
        c.sync()
        t0 = time.perf_counter()
        for it in range(20):
            _lib.check(lib.b2c_stencil_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None))
        c.sync()
        dt = (time.perf_counter() - t0) / 20
        print(w, h, n, "rb", rb, "us %.1f" % (dt * 1e6), flush=True)
    c.close()
