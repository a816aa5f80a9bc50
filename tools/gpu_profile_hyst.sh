#!/bin/bash
# --set full captures of the three hysteresis kernels of one config-2 step (after warm-up); tag $1
TAG=${1:-x}
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-extras"
ncu --set full --clock-control none --import-source on -k regex:k_uf_ -s 9 -c 3 -f -o gpurun_out/${TAG}_hyst $B > gpurun_out/${TAG}_ncu_hyst.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_hyst.log
