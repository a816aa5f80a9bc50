# quick GPU loop: GPU parity tests, default bench, 4K latency, one ncu capture of the stencil + launch list
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err; tail -3 gpurun_out/${TAG}_b.err
python bench.py --workload frame4k --steps 200 --warmup 10 --no-cpu > gpurun_out/${TAG}_b4k.json 2> gpurun_out/${TAG}_b4k.err; tail -3 gpurun_out/${TAG}_b4k.err
python - <<PY
import json
for f in ("${TAG}_b","${TAG}_b4k"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), "ms/step %.4f stencil %.4f frac %.3f hyst %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["hysteresis_ms"]), d.get("latency_ms"), "e2e", round(d["e2e"]["value"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/${TAG}_launches.csv | awk -F'","' 'NR>1 {print $5, $NF}' | tail -8
