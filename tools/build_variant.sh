#!/bin/bash
# Builds a variant of the library with extra nvcc flags into build/variants/libsvcm_<name>.so (A/B runs with
# tools/gpu_variants.py):   tools/build_variant.sh <name> "<extra nvcc flags>"
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/variant_$name
mkdir -p $obj $root/build/variants
cp -u $root/build/svcm/svcm_scene_builtin.o $root/build/svcm/svcm_bvh.o $obj/ 2>/dev/null || true
make -s -C $root/smallvcm_b200/csrc OBJDIR=$obj OUTDIR=$obj EXTRA_NVFLAGS="$flags" > /dev/null
mv $obj/libsvcm.so $root/build/variants/libsvcm_$name.so
grep -A1 "k_lightBounceILb0\|k_cameraBounceILb0\|k_connectILb0\|k_shadowILb0\|k_mergeLane" $obj/ptxas_kernels.log | grep -o "Function properties for [A-Za-z0-9_]*\|Used [0-9]* registers\|[0-9]* bytes spill stores" | paste - - - 2>/dev/null | sed 's/Function properties for _ZN4svcm//' | cut -c1-90
