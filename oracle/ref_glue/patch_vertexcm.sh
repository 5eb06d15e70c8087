#!/bin/sh
# TEST INFRASTRUCTURE. Usage: patch_vertexcm.sh <reference src dir> <out dir>
# Writes a throw-away copy of the reference's vertexcm.hxx into <out dir> with exactly two
# added lines that tell the (replaced) Rng which path is about to be traced:
#   mRng.SetPath(aIteration, 0, pathIdx);  before the light  sub-path of pathIdx (vertexcm.hxx:323)
#   mRng.SetPath(aIteration, 1, pathIdx);  before the camera sub-path of pathIdx (vertexcm.hxx:417)
# No arithmetic is touched. The copy lives in a temp dir for the duration of one compile.
set -e
src="$1"; out="$2"
sed -e 's/^\([[:space:]]*\)SubPathState lightState;/\1mRng.SetPath(aIteration, 0, pathIdx);\n&/' \
    -e 's/^\([[:space:]]*\)SubPathState cameraState;/\1mRng.SetPath(aIteration, 1, pathIdx);\n&/' \
    "$src/vertexcm.hxx" > "$out/vertexcm.hxx"
n=$(grep -c 'mRng.SetPath(aIteration' "$out/vertexcm.hxx")
if [ "$n" != "2" ]; then echo "patch_vertexcm.sh: expected 2 hooks, got $n" >&2; exit 1; fi
# everything else must be byte-identical
d=$(diff "$src/vertexcm.hxx" "$out/vertexcm.hxx" | grep -c '^[<>]' || true)
if [ "$d" != "2" ]; then echo "patch_vertexcm.sh: unexpected diff size $d" >&2; exit 1; fi
