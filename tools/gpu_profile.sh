#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain bench first (must exit 0), then the ncu launch list of the same command and
# one --set full capture of each hot kernel.  Outputs under gpurun_out/; tools/summarize_profiles.py turns them into
# the tracked summaries under profiles/.
set -x
TAG=${1:-r01}
B="python bench.py --steps 5 --warmup 3 --no-cpu"
$B > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err || { tail -5 gpurun_out/${TAG}_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 3 -c 1 -f -o gpurun_out/${TAG}_stencil $B > gpurun_out/${TAG}_ncu_stencil.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hysteresis -s 3 -c 1 -f -o gpurun_out/${TAG}_hyst $B > gpurun_out/${TAG}_ncu_hyst.log 2>&1
B4="python bench.py --workload frame4k --steps 5 --warmup 3 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_hysteresis -s 3 -c 1 -f -o gpurun_out/${TAG}_hyst4k $B4 > gpurun_out/${TAG}_ncu_hyst4k.log 2>&1
ls -la gpurun_out | tail -12
