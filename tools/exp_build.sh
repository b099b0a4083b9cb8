#!/bin/bash
# Timing experiments on the scan kernel OUTSIDE the product tree: copies the package to exp/pkg_<name>, applies a sed script to
# the kernel sources there (results become wrong on purpose), builds that copy's librr_maxcorr.so.  Nothing under exp/ is
# tracked or shipped.  EXP_EXTRA='-DNAME=value ...' adds compiler flags.  usage: tools/exp_build.sh NAME 'sed-script for rr_scan_umma.cu' ['sed-script for rr_device.cuh' ['sed-script for rr_plan.h']]
set -e
NAME=$1
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$ROOT/exp/pkg_$NAME
rm -rf "$D"; mkdir -p "$D"
cp -r "$ROOT/repeatresolver_b200" "$D/"
rm -rf "$D/repeatresolver_b200/csrc/build"
cp -r "$ROOT/include" "$D/include"
cd "$D/repeatresolver_b200/csrc"
sed -i 's#"../../include/#"'"$D"'/include/#' *.cu *.h *.cuh *.c *.cpp 2>/dev/null || true
[ -n "${2:-}" ] && sed -i "$2" rr_scan_umma.cu
[ -n "${3:-}" ] && sed -i "$3" rr_device.cuh
[ -n "${4:-}" ] && sed -i "$4" rr_plan.h
# experiments change the pair counts: no count check
sed -i 's#if ((int64_t)counters\[0\] != plan.part_pairs#if (false \&\& (int64_t)counters[0] != plan.part_pairs#' rr_abi.cu
make -j8 EXTRA="${EXP_EXTRA:-}" ../librr_maxcorr.so > /dev/null 2>&1 || { echo "build failed: $NAME"; make EXTRA="${EXP_EXTRA:-}" ../librr_maxcorr.so 2>&1 | tail -20; exit 1; }
echo "built $D"
