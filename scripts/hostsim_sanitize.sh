#!/bin/bash
# The lane-level CUDA source on the host (tests/hostsim) under AddressSanitizer + UndefinedBehaviorSanitizer: every golden check,
# recorded search and recorded placement with out-of-bounds accesses, misaligned loads, signed overflow ... trapped.
# Round 1: 131 cases, no report.  Usage: scripts/hostsim_sanitize.sh   (from the repo root; restores the plain library afterwards)
set -e
cd "$(dirname "$0")/.."
H=tests/hostsim
python -c "import sys; sys.path.insert(0,'tests'); import hostsim; hostsim.build()"
cp $H/libhostsim.so /tmp/libhostsim_plain.so
g++ -O1 -g -std=c++17 -ffp-contract=off -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas \
    -I $H/shim -I maple_b200/csrc $H/hostsim.cpp -o $H/libhostsim.so
trap 'cp /tmp/libhostsim_plain.so '$H'/libhostsim.so; touch '$H'/libhostsim.so' EXIT
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python -m pytest tests/test_kernel_source_host.py tests/test_place_scan_host.py -x -q -p no:cacheprovider
