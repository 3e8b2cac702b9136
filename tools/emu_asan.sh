#!/bin/sh
# The kernel sources (csrc/*.cuh) under AddressSanitizer + UndefinedBehaviorSanitizer: the CPU SIMT emulator build of
# tests/emu compiled with -fsanitize=address,undefined, driven by the emulator test files.  This is the stand-in for
# compute-sanitizer (closed on the GPU pool, profiles/r02_sanitizer_refusal.txt): out-of-bounds shared / global accesses,
# misaligned accesses, shifts and overflows that are undefined in C++ are caught in the same source lines the GPU runs.
# usage: tools/emu_asan.sh [log file]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LOG=${1:-$ROOT/profiles/r02_emu_asan.log}
SO=/tmp/libqb_emu_asan.so
g++ -std=c++20 -O1 -g -fPIC -shared -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -Wno-unused-function -Wno-unknown-pragmas -o $SO $ROOT/tests/emu/emu_main.cpp
cd $ROOT
{
  echo "g++ $(g++ -dumpversion) -fsanitize=address,undefined -fno-sanitize-recover=undefined, $(date -u +%F)"
  QB_EMU_SO=$SO LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest tests/test_emu_decode.py tests/test_emu_encode.py tests/test_emu_encode_ts.py -x -q -p no:cacheprovider 2>&1 | tail -15
} | tee $LOG
