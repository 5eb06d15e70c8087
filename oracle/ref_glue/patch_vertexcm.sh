#!/bin/sh
# TEST INFRASTRUCTURE. Usage: patch_vertexcm.sh <reference src dir> <out dir>
# Writes throw-away copies of the reference's vertexcm.hxx / pathtracer.hxx / eyelight.hxx into
# <out dir>; vertexcm.hxx with exactly two
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

# the two unidirectional integrators: one hook each, right after the pixel
# coordinates are derived (pathtracer.hxx:54-55, eyelight.hxx:58-59); phase 1 = camera stream
for f in pathtracer.hxx eyelight.hxx; do
    sed -e 's/^\([[:space:]]*\)const int y = pixID \/ resX;/&\n\1mRng.SetPath(aIteration, 1, pixID);/' \
        "$src/$f" > "$out/$f"
    n=$(grep -c 'mRng.SetPath(aIteration, 1, pixID)' "$out/$f")
    d=$(diff "$src/$f" "$out/$f" | grep -c '^[<>]' || true)
    if [ "$n" != "1" ] || [ "$d" != "1" ]; then echo "patch_vertexcm.sh: $f: hooks $n, diff $d" >&2; exit 1; fi
done
