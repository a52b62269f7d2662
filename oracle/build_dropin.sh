#!/bin/bash
# oracle/build_dropin.sh — TEST INFRASTRUCTURE.  Builds three executables into oracle/_ref/ (git-ignored):
#   templering_sfm_ref : the reference's own CLI, unmodified (cpp/src/templering_sfm.cpp as it lies)
#   templering_sfm_gpu : the same main() with the front-end definitions guarded out and host/sfmgpu_shim.hpp
#                        included instead — exactly the patch of INTEGRATION.md §2, applied to a scratch copy under
#                        /tmp (nothing of the reference is copied into the repository).
#   ate_keyframes      : the reference's own trajectory-evaluation tool (cpp/tools/ate_keyframes.cpp, unmodified): scores
#                        both runs against the ground truth of the synthetic ring dataset (tests/ring_dataset.py).
# Used by tests/test_gpu_dropin.py to compare the whole pipeline's outputs with and without the GPU front end.
set -e
REF_DIR=${REF_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
PKG=$HERE/../structure-from-motion-3d-reconstruction_b200
SRC=$REF_DIR/cpp/src/templering_sfm.cpp
[ -f "$SRC" ] || { echo "reference sources not present at $REF_DIR: keeping prebuilt binaries (if any)"; exit 0; }
mkdir -p "$HERE/_ref"
TMP=$(mktemp -d /tmp/sfm_dropin.XXXXXX)
# guard ranges (1-based, inclusive) of the pinned reference: sample_bilinear, Pyramid+build_pyr, shi_tomasi,
# LKConfig..KLTTracker, RelPose+find_E_ransac (replaced in place by sfmgpu_two_view.hpp, which calls the TU's own solver and SVD),
# global_desc_32+dot_desc; the front-end shim is included after the using-declarations (line 30)
awk '
  NR==31  { print "#ifdef USE_SFMGPU"; print "#include \"sfmgpu_shim.hpp\""; print "#endif" }
  NR==183 || NR==220 || NR==237 || NR==307 || NR==640 || NR==1100 { print "#ifndef USE_SFMGPU" }
  { print }
  NR==761 { print "#else"; print "#include \"sfmgpu_two_view.hpp\"  // RelPose + find_E_ransac on the GPU, with this TU own invert_K / eight_point_E / svd3" }
  NR==198 || NR==232 || NR==302 || NR==466 || NR==761 || NR==1129 { print "#endif" }
' "$SRC" > "$TMP/templering_sfm_dropin.cpp"
CXXF="-std=c++20 -O3 -DNDEBUG -w -I$REF_DIR/cpp/include"
g++ $CXXF "$SRC" -o "$HERE/_ref/templering_sfm_ref"
g++ $CXXF -pthread -ffp-contract=off -DUSE_SFMGPU -I"$PKG/host" "$TMP/templering_sfm_dropin.cpp" -o "$HERE/_ref/templering_sfm_gpu" \
    -L"$PKG" -lsfmgpu -Wl,-rpath,'$ORIGIN/../../structure-from-motion-3d-reconstruction_b200'
g++ $CXXF "$REF_DIR/cpp/tools/ate_keyframes.cpp" -o "$HERE/_ref/ate_keyframes"
rm -rf "$TMP"
echo "built $HERE/_ref/templering_sfm_ref, templering_sfm_gpu and ate_keyframes"
