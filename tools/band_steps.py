"""Back-to-back row-band steps (torchrun): per-step wall time, P2P vs NCCL rounds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from cudacam_b200 import bands, synth
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = H = 16384
y0, rows = bands.band_rows(H, world, rank)
band = synth.giga_rows(y0, y0 + rows, W, H)
for mode in ("p2p", "nccl"):
    be = bands.CudaBandBackend(W, rows, y0, H, device=local)
    be.load(band)
    if mode == "p2p":
        be.enable_p2p(dist, rank, world)
    bc = bands.BandCanny(be, rank, world, dist)
    for _ in range(3):
        bc.run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ts = []
    t0 = time.perf_counter()
    for i in range(8):
        r = bc.run()
        torch.cuda.synchronize()
        t1 = time.perf_counter(); ts.append(1e6 * (t1 - t0)); t0 = t1
    if rank == 0:
        print(mode, "rounds", r, "step us:", " ".join("%.0f" % t for t in ts), flush=True)
    dist.barrier()
    be.close()
dist.destroy_process_group()
