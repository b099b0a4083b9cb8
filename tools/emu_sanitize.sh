#!/bin/bash
# The kernel-emulation tests (tests/test_kernel_emulation.py) with the emulation driver built under a sanitizer.
#   bash tools/emu_sanitize.sh address   every global and static-shared-memory access of the emulated kernels (packing,
#                                        AND+POPC scan, Cliquer 1-3, Relative_Vars, Kmeans) is bounds-checked
#   bash tools/emu_sanitize.sh thread    the CUDA threads are real pthreads and the barriers real barriers, so a missing
#                                        __syncthreads()/__syncwarp() between a write and a read of shared or global memory
#                                        shows up as a data race.  Expected reports: only the byte flags mark[] of
#                                        rr_k_relvars_pairs (racy by design: only ever set to 1, read as a hint)
# Test infrastructure; three to four minutes each.
set -eu
mode="${1:-address}"
cd "$(dirname "$0")/.."
mkdir -p tests/emu/_build
lib="tests/emu/_build/libemu_${mode}.so"
g++ -O1 -g -std=c++17 -fPIC -shared -Wl,-Bsymbolic -ffp-contract=off -fsanitize="$mode" -fno-omit-frame-pointer \
    -Itests/emu -Irepeatresolver_b200/csrc -o "$lib" tests/emu/emu_driver.cpp -lpthread
log="/tmp/rr_emu_${mode}"
rm -f "$log".*
if [ "$mode" = thread ]; then
    export LD_PRELOAD="$(g++ -print-file-name=libtsan.so)" TSAN_OPTIONS="halt_on_error=0:report_signal_unsafe=0:log_path=$log"
else
    export LD_PRELOAD="$(g++ -print-file-name=libasan.so)" ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:log_path=$log"
fi
RR_EMU_LIB="$PWD/$lib" python -m pytest tests/test_kernel_emulation.py -q -p no:cacheprovider
unset LD_PRELOAD
if ls "$log".* >/dev/null 2>&1; then
    echo "sanitizer reports (top frames):"
    grep -h "^    #0" "$log".* | grep -v "pthread_create\|malloc" | sed 's/(lib.*//' | sort | uniq -c | sort -rn | head -20
    [ "$mode" = thread ] || exit 1
else
    echo "no sanitizer report"
fi
