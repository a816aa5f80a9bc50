#!/bin/bash
# ncu --set full capture of the stencil kernel of a profiling build: gpu_profile_variant.sh TAG VARIANT
TAG=$1; V=$2
B2C_LIB_PATH=$PWD/build_variants/x$V.so ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 3 -c 1 -f -o gpurun_out/${TAG}_x$V python tools/sweep_rb.py 272 > gpurun_out/${TAG}_x$V.log 2>&1
tail -2 gpurun_out/${TAG}_x$V.log
