#!/bin/bash
# The kernel-emulation tests (tests/test_kernel_emulation.py) with the emulation driver built under AddressSanitizer: every
# global and static-shared-memory access of the emulated kernels (packing, AND+POPC scan, Cliquer 1-3, Relative_Vars,
# Kmeans) is bounds-checked.  Test infrastructure; about three minutes.  Usage: bash tools/emu_asan.sh
set -eu
cd "$(dirname "$0")/.."
mkdir -p tests/emu/_build
g++ -O1 -g -std=c++17 -fPIC -shared -Wl,-Bsymbolic -ffp-contract=off -fsanitize=address -fno-omit-frame-pointer \
    -Itests/emu -Irepeatresolver_b200/csrc -o tests/emu/_build/libemu_asan.so tests/emu/emu_driver.cpp -lpthread
rm -f /tmp/rr_emu_asan.*
LD_PRELOAD="$(g++ -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:log_path=/tmp/rr_emu_asan \
    RR_EMU_LIB="$PWD/tests/emu/_build/libemu_asan.so" python -m pytest tests/test_kernel_emulation.py -q -p no:cacheprovider
if ls /tmp/rr_emu_asan.* >/dev/null 2>&1; then echo "AddressSanitizer reports:"; head -20 /tmp/rr_emu_asan.*; exit 1; fi
echo "no AddressSanitizer report"
