#!/bin/bash
# Build a second copy of libsfno_b200.so with engine experiment flags, next to the shipped one (run in the build
# container; the result travels to the GPU box with the snapshot):
#   scripts/build_variant.sh ldtm_pair -DSFNO_TC_LDTM_PAIR=1
#   -> build/ldtm_pair/pkg/libsfno_b200.so      use it with  SFNO_B200_LIB=build/ldtm_pair/pkg/libsfno_b200.so
set -eu
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=$ROOT/build/$NAME
rm -rf "$V"; mkdir -p "$V/pkg/csrc" "$V/include"
cp "$ROOT"/spherical-dyffusion_b200/csrc/{*.cu,*.cuh,*.cpp,*.h,Makefile} "$V/pkg/csrc/"
cp "$ROOT"/include/*.h "$V/include/"
make -C "$V/pkg/csrc" -j"$(nproc)" EXPERIMENT_FLAGS="$*" > "$V/build.log" 2>&1 || { tail -20 "$V/build.log"; exit 1; }
grep -h "spill" "$V"/pkg/csrc/*.ptxas.log | sort | uniq -c | sort -rn | head -5
ls -la "$V/pkg/libsfno_b200.so"
