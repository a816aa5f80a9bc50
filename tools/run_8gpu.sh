# ON AN 8-GPU BOX: multi-rank parity, giga strong scaling at 2, 4 and 8 ranks with phases, batch line at 8 ranks
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
run() { g=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for g in 2 4 8; do
  run $g bench.py --gpus $g --workload giga --steps 20 --warmup 5 > gpurun_out/scale_giga_n$g.json 2> gpurun_out/scale_giga_n$g.err || tail -3 gpurun_out/scale_giga_n$g.err
done
B2C_NO_SAMPLER=1 run 8 bench.py --gpus 8 --workload giga --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('giga n8 without the clock sampler: ms/step', d['ms_per_step'])"
run 4 tools/band_phases.py 2>&1 | grep "p2p\] rank" | sort | head -4
run 8 tools/band_phases.py 2>&1 | grep "p2p\] rank" | sort
run 8 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_batch_n8.json 2> gpurun_out/scale_batch_n8.err || tail -3 gpurun_out/scale_batch_n8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*_n[248].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        g = d.get("giga") or {}
        print(f.split("/")[-1], "n", d["n_gpus"], "value %.0f ms/step %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), g.get("ms_per_step"), g.get("equals_oracle_golden"), g.get("phase_us_rank0"))
    except Exception as e:
        print(f, "unreadable", e)
PY
