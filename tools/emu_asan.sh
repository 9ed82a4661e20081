#!/bin/bash
# Memory check of the product's kernels WITHOUT a GPU: the emulated library (tests/emu) built with AddressSanitizer, then the emulated
# suite, the -m gpu parity files (against the emulated context) and a slice of the fuzzer run on it.  Device buffers are heap blocks,
# __shared__ arrays are thread-local globals, local arrays live on the fibers' stacks: an out-of-bounds access of a kernel is reported
# like compute-sanitizer's memcheck would.  TEST INFRASTRUCTURE ONLY.
#
#     tools/emu_asan.sh [fuzz seconds, default 300]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=/tmp/liblinevis_b200_emu_asan.so
python "$ROOT/tests/emu/build_emu.py" > /dev/null          # regenerates tests/emu/_gen
(cd "$ROOT/tests/emu/_gen" && g++ -O1 -g -std=c++17 -fopenmp -DLV_HOST_EMU -I/usr/local/cuda/include -I"$ROOT/linevis_b200/csrc" \
    -ffp-contract=off -fno-fast-math -march=x86-64-v3 -fPIC -shared -Wl,-Bsymbolic -Wno-attributes -Wno-unknown-pragmas \
    -Wno-subobject-linkage -fsanitize=address -fno-omit-frame-pointer -o "$OUT" emu_main.cpp)
export LD_PRELOAD="$(gcc -print-file-name=libasan.so)"
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1
export LV_EMU_LIB="$OUT"
cd "$ROOT"
python -m pytest tests/test_emu_simt.py -x -q -k "not fuzz and not nan_rays"
LV_EMU_CTX=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prebaker.py tests/test_gpu_tritubes.py -x -q -m gpu \
    -k "not pack_unpack and not owned_tiles and not tile_sharding and not peer_frame and not frame_to_rgba8"   # those need torch.cuda
python tools/fuzz_emu.py --seconds "${1:-300}" --seed 31
