"""Row-band step by phase (torchrun, one rank per GPU): BandCanny.phase_times + the kernel-level marks of the band
hysteresis / seam pass ("seam_phase_us*").  usage: torchrun ... tools/band_phases.py [W H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from cudacam_b200 import bands, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
y0, rows = bands.band_rows(H, world, rank)
be = bands.CudaBandBackend(W, rows, y0, H, device=local)
be.load(synth.giga_rows(y0, y0 + rows, W, H))
for mode in (["p2p", "nccl"] if world > 1 else ["single"]):
    if mode == "p2p":
        be.enable_p2p(dist, rank, world)
    else:
        be.p2p = False
    bc = bands.BandCanny(be, rank, world, dist if world > 1 else None)
    for _ in range(5):
        bc.run()
    be.sync()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        bc.run()
    b.record()
    be.sync()
    step = a.elapsed_time(b) * 100
    ph = bc.phase_times()
    be.set_phase_timing(True)
    for _ in range(3):
        bc.run()
    be.sync()
    k = be.seam_phase_us()
    be.set_phase_timing(False)
    print(f"[{mode}] rank {rank} step {step:.0f} us | " + " ".join(f"{n}={v:.0f}" for n, v in ph.items()) + " | " + " ".join(f"{n}={v}" for n, v in k.items()), flush=True)
    if world > 1:
        dist.barrier()
be.close()
if world > 1:
    dist.destroy_process_group()
