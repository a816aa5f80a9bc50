#!/bin/bash
# quick: one --set full capture of the stencil kernel of the default bench
TAG=${1:-x}
B="python bench.py --steps 3 --warmup 3 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 3 -c 1 -f -o gpurun_out/${TAG}_stencil $B > gpurun_out/${TAG}_ncu_stencil.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_stencil.log
