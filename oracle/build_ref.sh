#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/libsphref.so from the reference sources WHERE THEY LIE under
# $SPH_REFERENCE (default /root/reference).  Outputs go only into oracle/_ref/ (git-ignored).
# The reference's own build system (gcc-8 + CUDA 6.5 + GL, source/CMakeLists.txt) is not used
# and cannot be: no GL headers here, and texture references no longer exist in CUDA 12.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SPH_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
GEN="$OUT/gen"
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"

if [ ! -f "$REF/source/CUDA/System.cu" ]; then
    echo "build_ref: no reference at $REF (fine on the GPU box: the prebuilt .so travels)"; exit 0
fi
mkdir -p "$GEN"

# 1. kernel text: the block System.cu itself marks as the old Kernel.cu (System.cu:9,548)
sed -n '11,548p' "$REF/source/CUDA/System.cu" > "$GEN/ref_kernels.inc"

# 2. host-side scene/initialiser text that cannot be included verbatim because its
#    includes drag in GL (pch/header.h, App/App.h): strip the #include lines only.
sed -n '32,46p'   "$REF/source/pch/header.h" > "$GEN/header_ops.inc"
sed -n '123,156p' "$REF/source/pch/header.h" > "$GEN/header_util.inc"
for f in Scene.cpp Scene_Load.cpp SPH_Init.cpp SPH_Scenes.cpp; do
    grep -v '^[[:space:]]*#include' "$REF/source/SPH/$f" > "$GEN/${f%.cpp}.inc"
done
# per-step host prologue (App::UpdateEmitter), for the "next" row N1
grep -v '^[[:space:]]*#include' "$REF/source/App/Update.cpp" > "$GEN/App_Update.inc"
sed -n '588,593p' "$REF/source/App/Input.cpp" > "$GEN/App_mulTr.inc"

CXXFLAGS="-O2 -fopenmp -fPIC -std=c++14 -ffp-contract=off -w"
INC="-I$HERE -I$GEN -I$CUDA_INC -I$REF/source/CUDA -I$REF/source/external -I$REF/source/external/cutil"

g++ $CXXFLAGS $INC -c "$HERE/ref_driver.cpp" -o "$OUT/ref_driver.o"
if [ -f "$HERE/ref_host.cpp" ]; then
    g++ $CXXFLAGS $INC -I"$REF/source/external/tinyxml" -I"$REF/source" \
        -c "$HERE/ref_host.cpp" -o "$OUT/ref_host.o"
    for t in tinyxml tinystr tinyxmlerror tinyxmlparser; do
        g++ -O2 -fPIC -w -c "$REF/source/external/tinyxml/$t.cpp" -o "$OUT/$t.o"
    done
    g++ -shared -fopenmp -o "$OUT/libsphref.so" "$OUT"/ref_driver.o "$OUT"/ref_host.o \
        "$OUT"/tinyxml.o "$OUT"/tinystr.o "$OUT"/tinyxmlerror.o "$OUT"/tinyxmlparser.o
else
    g++ -shared -fopenmp -o "$OUT/libsphref.so" "$OUT"/ref_driver.o
fi
echo "build_ref: built $OUT/libsphref.so"
