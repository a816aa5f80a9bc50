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
for K in k_uf_tile k_uf_border k_uf_resolve; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_$K $B > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out | tail -12
