#!/bin/bash
# one --set full capture (with source) of a kernel (regex $2) of a bench workload ($3, default batch1080p); tag $1
TAG=${1:-x}; K=${2:-k_stencil}; WL=${3:-batch1080p}
B="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-extras"
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_${K} $B > gpurun_out/${TAG}_ncu_${K}.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_${K}.log
