"""Debug helper: the three stencil kernels (0 marching, 1 tile, 2 fused CTA-tile) on the GPU against the oracle; prints
where the 2-bit maps differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cudacam_b200 as cb
import oracle_py as O


def dec(m, w):
    s = np.zeros((m.shape[0], m.shape[1] * 16), np.uint8)
    for b in range(16):
        s[:, b::16] = ((m >> b) & 1) * 255 + ((m >> (16 + b)) & 1) * 128
    return s[:, :w]


for kind, w, h, seed in [("scene", 240, 60, 1), ("scene", 1280, 720, 0xC0FFEE), ("noise", 480, 120, 2), ("steps", 480, 120, 3)]:
    f = cb.synth.frame(kind, seed, w, h)
    want = O.canny(f, want_edges=False)["thresh"]
    with cb.CannyEdge(w, h) as c:
        c.run(f)
        got = dec(c.map2(), w)
        c.set_option("stencil_impl", 1)
        c.run(f)
        tile = dec(c.map2(), w)
        c.set_option("stencil_impl", 2)
        c.run(f)
        fused = dec(c.map2(), w)
    d = np.argwhere(got != want)
    print(kind, w, h, "tile ok", np.array_equal(tile, want), "fused ok", np.array_equal(fused, want), "march diffs", len(d))
    if len(d):
        print("  rows", d[:, 0].min(), d[:, 0].max(), "cols", d[:, 1].min(), d[:, 1].max())
        print("  got/want histogram:", {(int(a), int(b)): int(((got == a) & (want == b)).sum()) for a in (0, 128, 255) for b in (0, 128, 255) if a != b})
        print("  row%60 hist", np.bincount(d[:, 0] % 60, minlength=60))
        print("  col%240%8 hist", np.bincount((d[:, 1] % 240) % 8, minlength=8))
        print("  first", d[:12].tolist())
