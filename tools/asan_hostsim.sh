#!/bin/sh
# Memory-safety net for the device headers (compute-sanitizer is closed on the GPU pool): the CPU lane-emulation
# harness -- the same PSI_HD code the kernels run -- built with AddressSanitizer + UBSan and driven by the CPU tests
# that exercise it (quadrature, AAA scratch carving, PSIMode searches, refinement sweep, terminal-velocity map).
#     sh tools/asan_hostsim.sh [pytest args]
set -e
cd "$(dirname "$0")/.."
mkdir -p _exp
g++ -O1 -g -std=c++17 -shared -fPIC -ffp-contract=off -fopenmp -fsanitize=address,undefined -fno-omit-frame-pointer \
    -x c++ -o _exp/libpsi_hostsim_asan.so tests/hostsim/hostsim.cpp -ldl -lpthread
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1
export PSI_HOSTSIM_LIB="$PWD/_exp/libpsi_hostsim_asan.so"
if [ $# -eq 0 ]; then set -- tests/test_hostsim.py tests/test_grid_refine.py tests/test_guessgen.py tests/test_tables.py; fi
python -m pytest -x -q -m "not gpu" -p no:cacheprovider "$@"
