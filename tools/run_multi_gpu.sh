# Runs ON A MULTI-GPU BOX (gpurun --gpus N): NCCL row-band tests, then the scaling lines of the three partitioned workloads.
# usage: run_multi_gpu.sh N [workloads...]   (default: batch giga streams)
N=${1:-2}; shift
WLS=${@:-batch giga streams}
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
run() { # $1 = gpus, rest = bench args
  g=$1; shift
  if [ "$g" = 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g "$@"; fi
}
for g in ${GS:-1 2 4 8}; do
  [ $g -le $N ] || continue
  for wl in $WLS; do
    case $wl in
      batch) run $g --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_batch_n$g.json 2> gpurun_out/scale_batch_n$g.err || tail -3 gpurun_out/scale_batch_n$g.err;;
      giga) run $g --workload giga --steps 5 --warmup 3 > gpurun_out/scale_giga_n$g.json 2> gpurun_out/scale_giga_n$g.err || tail -3 gpurun_out/scale_giga_n$g.err;;
      streams) run $g --workload streams1080p --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_streams_n$g.json 2> gpurun_out/scale_streams_n$g.err || tail -3 gpurun_out/scale_streams_n$g.err;;
    esac
  done
done
if [ $N -gt 1 ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/band_phases.py 2>&1 | grep "rank"; fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n", d["n_gpus"], "value %.0f Mpx/s  ms/step %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), (d.get("giga") or {}).get("phase_us_rank0", ""), (d.get("giga") or {}).get("equals_oracle_golden", ""))
    except Exception as e:
        print(f, "unreadable", e)
PY
