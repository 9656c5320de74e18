#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/OpticalFlow_ref (and SampleTextureToVertices_ref): the reference's own OpticalFlow.cpp + include/ headers,
# compiled from where they lie under $MOF_REFERENCE (default /root/reference) with the shipped
# Makefile's release flags (OpticalFlow/Makefile:6,12), MKL off as shipped (OpticalFlow.cpp:29).
# No reference source is copied into the repository: the few generated files below live in a
# temporary directory that is deleted after the compile; only the binary lands in oracle/_ref/.
#
# What has to be generated, and why (the sources are MSVC-dialect; the shipped Makefile links
# GL/GLU/glut/GLEW/libpng, none of which exist in this image):
#   1. `Misha\Image.h`   — Src/VectorIO.h:6 includes a path with a backslash; a forwarding header.
#   2. Src/{VectorField.h,Whitney.inl,Conformal.inl,Connection.inl} — the derived classes use base
#      members (coeffs, prolongationOperator, ...) unqualified, which ISO two-phase lookup rejects;
#      `using VectorField<Real>::...;` lines are added by sed. VectorField.h is copied beside them
#      unmodified so that its quoted includes (:114-116) resolve to the patched copies.
#   3. OpticalFlow.gen.cpp — OpticalFlow.cpp with the undeclared name `eFlowField` (:308,:323, in
#      member functions that are never instantiated) replaced by `tFlowField`.
#   4. shims/ (committed, our own code): headless stand-ins for <GL/glew.h>, <GL/glut.h>,
#      <Src/SurfaceVisualization.inl> and a zlib-backed <Misha/PNG.h>.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${MOF_REFERENCE:-/root/reference}"
OUT="$REPO/oracle/_ref"
if [ ! -f "$REF/OpticalFlow/OpticalFlow.cpp" ]; then
    echo "build_ref: no reference at $REF (expected on the GPU box: the prebuilt binary travels instead)"
    exit 0
fi
mkdir -p "$OUT"
GEN="$(mktemp -d "$OUT/gen.XXXXXX")"
trap 'rm -rf "$GEN"' EXIT
mkdir -p "$GEN/Src"

printf '#include <Misha/Image.h>\n' > "$GEN/Misha\\Image.h"
cp "$REF/include/Src/VectorField.h" "$GEN/Src/VectorField.h"
USING='\tusing VectorField<Real>::coeffs; using VectorField<Real>::prolongationOperator; using VectorField<Real>::restrictionOperator; using VectorField<Real>::smoothOperator;'
for f in Whitney Conformal Connection; do
    sed -E "0,/^public:/s//public:\n$USING/" "$REF/include/Src/$f.inl" > "$GEN/Src/$f.inl"
done
sed -e 's/eFlowField/tFlowField/g' "$REF/OpticalFlow/OpticalFlow.cpp" > "$GEN/OpticalFlow.gen.cpp"

CXXFLAGS="-fpermissive -fopenmp -Wno-deprecated -Wno-unused-result -Wno-format -msse2 -std=c++14 -O3 -DRELEASE -funroll-loops -ffast-math -DNDEBUG -w"
g++ $CXXFLAGS \
    -I"$GEN" -I"$HERE/shims" -I"$REPO/meshopticalflow_b200/csrc/host" -I"$REF/include" \
    "$HERE/ref_driver.cpp" "$REPO/meshopticalflow_b200/csrc/host/png_codec.cpp" \
    -o "$OUT/OpticalFlow_ref" -lgomp -lz
echo "build_ref: built $OUT/OpticalFlow_ref"
# The sibling tool that makes per-vertex inputs from the texture configuration's files; compiles as it is.
g++ $CXXFLAGS -I"$HERE/shims" -I"$REPO/meshopticalflow_b200/csrc/host" -I"$REF/include" \
    "$REF/SampleTextureToVertices/SampleTextureToVertices.cpp" "$REPO/meshopticalflow_b200/csrc/host/png_codec.cpp" \
    -o "$OUT/SampleTextureToVertices_ref" -lgomp -lz
echo "build_ref: built $OUT/SampleTextureToVertices_ref"
