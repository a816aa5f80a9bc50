"""Golden hash of BASELINE configs[4]: the synthetic 16384x16384 mosaic (cudacam_b200.synth.giga_rows) through the CPU
oracle (oracle/canny_oracle.c, pinned against the reference's own kernels by tests/golden/*.npz), sha256 of the u8 edge
map.  TEST INFRASTRUCTURE: run on the CPU (about a minute, ~6 GB of RAM); writes tests/golden/giga_sha256.json, which
bench.py's `giga` record and tests/test_gpu_multi.py compare the sharded GPU result with.

    python oracle/make_giga_golden.py [W H]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O   # noqa: E402
from cudacam_b200 import synth   # noqa: E402  (host-side frame generator only)

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
t0 = time.time()
img = synth.giga_rows(0, H, W, H)
t1 = time.time()
edges = O.canny(img)["edges"]
t2 = time.time()
digest = hashlib.sha256(np.ascontiguousarray(edges).tobytes()).hexdigest()
path = os.path.join(ROOT, "tests", "golden", "giga_sha256.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[f"{W}x{H}"] = digest
d[f"{W}x{H}_edge_pixels"] = int((edges == 255).sum())
d["how"] = "oracle/make_giga_golden.py: synth.giga_rows -> oracle_canny (thresholds 10/40) -> sha256 of the u8 edge map"
json.dump(d, open(path, "w"), indent=1)
print(f"{W}x{H}: {digest}  edge pixels {d[f'{W}x{H}_edge_pixels']}  (generate {t1 - t0:.1f} s, oracle {t2 - t1:.1f} s)")
