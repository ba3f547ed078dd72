#!/bin/bash
# ThreadSanitizer over the kernel sources under host emulation (tests/emu): every lane is a host thread and __syncwarp /
# __syncthreads / the shuffles are barriers, so an unsynchronised shared-memory access between lanes - a race on the GPU
# too - is a data race TSan reports.  (compute-sanitizer's racecheck is not available on the GPU pool.)
#   tools/emu_tsan.sh [pytest -k expression]
set -e
cd "$(dirname "$0")/.."
OUT=/tmp/libzkb_emu_tsan.so
g++ -std=c++20 -O1 -g -fsanitize=thread -fPIC -shared -pthread -o $OUT tests/emu/emu_kernels.cpp
TSAN=$(g++ -print-file-name=libtsan.so)
LD_PRELOAD=$TSAN ZKB_EMU_LIB=$OUT TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0 log_path=/tmp/zkb_tsan" \
  python -m pytest tests/test_emu_kernels.py -x -q -k "${1:-squaring}" -p no:cacheprovider
echo "TSan reports:"; ls /tmp/zkb_tsan.* 2>/dev/null | wc -l
