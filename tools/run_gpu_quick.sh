# quick GPU loop: GPU parity tests, default bench (no extras), one ncu counter pass of the stencil + launch list
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err; tail -3 gpurun_out/${TAG}_b.err
python - <<PY
import json
for f in ("${TAG}_b",):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), "ms/step %.4f stencil %.4f frac %.3f hyst %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["hysteresis_ms"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("pcie_ceiling_rank0"))
PY
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers --clock-control none -k regex:k_stencil -s 3 -c 1 --csv --log-file gpurun_out/${TAG}_stencil_counters.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/${TAG}_stencil_counters.csv | awk -F'","' 'NR>1 {print $(NF-2), $NF}' | tr -d '"'
