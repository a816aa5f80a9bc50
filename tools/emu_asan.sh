#!/bin/bash
# The kernels' indexing under AddressSanitizer: the CPU emulation of the .cuh sources (tests/emu) built with
# -fsanitize=address, the emulator tests run with libasan preloaded.  (compute-sanitizer is closed on the GPU pool; this
# is how the out-of-bounds read of the tile kernel below a row band was confirmed fixed.)
set -e
cd "$(dirname "$0")/.."
g++ -std=c++20 -O1 -g -fsanitize=address -fno-omit-frame-pointer -pthread -fPIC -shared -fvisibility=hidden -Wno-unused -DB2C_EMU_FUSED -I tests/emu -o /tmp/libemu_asan.so tests/emu/emu_main.cpp
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) B2C_EMU_SO=/tmp/libemu_asan.so python -m pytest tests/test_emu_kernels.py tests/test_bands_gloo.py -x -q -p no:cacheprovider -k "not sharded_image and not one_exchange" "$@"   # (the two deselected tests spawn gloo workers)
