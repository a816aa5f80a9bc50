"""Times the marching stencil kernel at reduced occupancy (extra dynamic shared memory per CTA) -- GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cudacam_b200 as cb
from cudacam_b200 import _lib
lib = _lib.lib
w, h, n = 1920, 1080, 64
host = cb.synth.batch("scene", n, w, h, distinct=8)
d_in = torch.from_numpy(host.reshape(-1)).cuda()
c = cb.CannyEdge(w, h, max_batch=n)
for rb in (44, 92, 272):
  for extra in (0, 1100, 2400, 3900, 5700, 7800, 12000, 20000):
    c.set_option("march_rb", rb); c.set_option("march_extra_smem", extra)
    for it in range(3): _lib.check(lib.b2c_stencil_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None))
    c.sync(); t0 = time.perf_counter()
    for it in range(10): _lib.check(lib.b2c_stencil_device(c._h, d_in.data_ptr(), w * 3, w * 3 * h, n, None))
    c.sync()
    ctas = 233472 // (15504 + extra + 1024)
    print("rb", rb, "extra", extra, "ctas/sm<=", min(14, ctas), "us %.1f" % ((time.perf_counter() - t0) / 10 * 1e6), flush=True)
