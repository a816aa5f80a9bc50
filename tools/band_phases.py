"""Per-phase wall times of one row-band step (torchrun, N ranks): where the time of BASELINE config 5 goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from cudacam_b200 import bands, synth, _lib
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = H = 16384
y0, rows = bands.band_rows(H, world, rank)
be = bands.CudaBandBackend(W, rows, y0, H, device=local)
be.load(synth.giga_rows(y0, y0 + rows, W, H))
if world > 1 and os.environ.get("B2C_BAND_NCCL", "0") != "1":
    be.enable_p2p(dist, rank, world)
bc = bands.BandCanny(be, rank, world, dist if world > 1 else None)
for _ in range(3):
    bc.run()
def T():
    torch.cuda.synchronize()
    return time.perf_counter()
for rep in range(2):
    if world > 1: dist.barrier()
    t = [T()]; names = []
    be.ghost(0).zero_(); be.ghost(1).zero_()
    bc.exchange_input_halos(); t.append(T()); names.append("halo")
    be.stencil(); t.append(T()); names.append("stencil")
    be.hysteresis(True, write_edges=True); t.append(T()); names.append("hyst0")
    r = 1
    if world > 1 and getattr(be, "p2p", False):
        r = be.converge(); t.append(T()); names.append("p2p_rounds")
    while world > 1 and not getattr(be, "p2p", False):
        bc._exchange(be.boundary(0), be.ghost(0), be.boundary(1), be.ghost(1)); t.append(T()); names.append("xchg")
        be.hysteresis(False, write_edges=True); t.append(T()); names.append("reentry")
        flag = be.seeded().clone(); dist.all_reduce(flag, op=dist.ReduceOp.MAX); v = int(flag.item()); t.append(T()); names.append("allreduce")
        if v == 0: break
        r += 1
    t.append(T()); names.append("-")
    if getattr(be, "p2p", False):
        n = _lib.lib.b2c_get_info(be._h, b"p2p_stamp55")
        st = [_lib.lib.b2c_get_info(be._h, b"p2p_stamp%d" % k) & 0xFFFFFFFF for k in range(n)]
        print("rank", rank, "rounds kernel stamps (us since start):", " ".join("%.1f" % (((x - st[0]) & 0xFFFFFFFF) / 1000.0) for x in st), flush=True)
    if rank == 0:
        print("rep", rep, "rounds", r, " ".join(f"{n}={1e6*(b-a):.0f}us" for n, a, b in zip(names, t, t[1:])), "total=%.0fus" % (1e6 * (t[-1] - t[0])))
be.close()
if world > 1:
    dist.destroy_process_group()
