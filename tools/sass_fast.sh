#!/bin/bash
# SASS of the fast-path instantiation of the scan kernel (rr_k_scan_umma<1,2,0,1>) of a built library, without the encoding
# lines: tools/sass_fast.sh [path/to/librr_maxcorr.so] > out.sass
SO=${1:-$(dirname "$0")/../repeatresolver_b200/librr_maxcorr.so}
cuobjdump -sass "$SO" | awk '/Function : .*rr_k_scan_ummaILb1ELi2ELb0ELb1/ {on=1} on && /Function : / && !/rr_k_scan_ummaILb1ELi2ELb0ELb1/ {on=0} on' \
  | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's/[ \t]*$//'
