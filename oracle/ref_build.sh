#!/bin/bash
# ORACLE — TEST INFRASTRUCTURE ONLY. Builds oracle/_ref/ from the reference's OWN sources, compiled where they lie
# under /root/reference (nothing is copied into the repo), against the third-party stand-ins in oracle/ref_shim/
# (OpenCV, Eigen, gflags, glog, folly are absent from this image; the OpenCV image primitives forward to the
# cv2-pinned oracle/ocv.h). Outputs (git-ignored, shipped to the GPU box with the snapshot):
#   oracle/_ref/TestRenderStereoPanorama   the reference's frame renderer (its own main(), flags, threading, file I/O)
#   oracle/_ref/libref.so                  the reference's operators behind a small C API (oracle/ref_capi.cpp)
#   oracle/_ref/libref_preview.so, libref_isp.so   TestHyperPreview and the soft ISP (camera_isp/CameraIsp.h) behind C APIs
#   oracle/_ref/libref_unpack.so           the footage container reader and raw-sample converter of the Unpacker (BinaryFootageFile, RawConverter)
# Flags follow the reference's build (surround360_render/CMakeLists.txt: -O3 -std=c++11 -mavx, no -mfma, threads).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${S360_REFERENCE_ROOT:-/root/reference}/surround360_render/source
if [ ! -d "$REF" ]; then echo "ref_build: $REF not found (prebuilt oracle/_ref is used as is)"; exit 0; fi
OUT="$HERE/_ref"
mkdir -p "$OUT/obj"
CXX=${REF_CXX:-/usr/bin/g++}   # not $CXX: the image presets it to a toolchain that links libstdc++ statically (two C++ runtimes in one python process)
CXXFLAGS="-m64 -O3 -mmmx -mavx -msse -msse3 -funroll-loops -pipe -std=c++14 -pthread -fPIC -w -g"
INC="-I$HERE/ref_shim -I$REF/util -I$REF/render -I$REF/optical_flow -I$REF/calibration -I$REF"
SRCS="util/CvUtil.cpp util/SystemUtil.cpp util/StringUtil.cpp render/Camera.cpp render/ImageWarper.cpp render/PoleRemoval.cpp render/RigDescription.cpp optical_flow/NovelView.cpp"
OBJS=""
pids=""
for s in $SRCS; do
  o="$OUT/obj/$(echo $s | tr '/' '_' | sed 's/\.cpp$/.o/')"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ] || [ -n "$(find "$HERE/ref_shim" "$HERE/ocv.h" -newer "$o" -print -quit)" ]; then
    $CXX $CXXFLAGS $INC -c "$REF/$s" -o "$o" &
    pids="$pids $!"
  fi
done
o="$OUT/obj/ref_shim.o"; OBJS="$OBJS $o"
$CXX $CXXFLAGS $INC -c "$HERE/ref_shim/ref_shim.cpp" -o "$o" &
pids="$pids $!"
$CXX $CXXFLAGS $INC -c "$REF/test/TestRenderStereoPanorama.cpp" -o "$OUT/obj/trsp_main.o" &
pids="$pids $!"
$CXX $CXXFLAGS $INC -c "$HERE/ref_capi.cpp" -o "$OUT/obj/ref_capi.o" &
pids="$pids $!"
for p in $pids; do wait $p; done
$CXX -pthread -o "$OUT/TestRenderStereoPanorama" "$OUT/obj/trsp_main.o" $OBJS -lz
# The reference's renderer with the B200 flow adapter of integration/reference_tree/optical_flow/B200Flow.h selectable through
# --side_flow_alg / --polar_flow_alg b200_pixflow_low (the translation units that call the flow factory are compiled again with
# the factory hook pre-included; they link the product library like a maintainer's build would).
REPO="$(cd "$HERE/.." && pwd)"
if [ -f "$REPO/surround360_b200/libs360b200.so" ]; then
  HOOK="-include $HERE/ref_shim/b200_factory_hook.h -I$REPO/integration/reference_tree/optical_flow -I$REPO/include"
  B200OBJS=""
  for s in test/TestRenderStereoPanorama.cpp optical_flow/NovelView.cpp render/PoleRemoval.cpp; do
    o="$OUT/obj/b200_$(echo $s | tr '/' '_' | sed 's/\.cpp$/.o/')"
    $CXX $CXXFLAGS $INC $HOOK -c "$REF/$s" -o "$o" &
    B200OBJS="$B200OBJS $o"
  done
  wait
  REST=$(for o in $OBJS; do case "$o" in *optical_flow_NovelView.o|*render_PoleRemoval.o) ;; *) echo -n "$o ";; esac; done)
  $CXX -pthread -o "$OUT/TestRenderStereoPanorama_b200" $B200OBJS $REST -lz -L"$REPO/surround360_b200" -ls360b200 -Wl,-rpath,'$ORIGIN/../../surround360_b200'
fi
$CXX -pthread -shared -o "$OUT/libref.so" "$OUT/obj/ref_capi.o" $OBJS -lz
# the reference's TestHyperPreview behind its own small C API (its gflags clash with the renderer's: separate library)
$CXX $CXXFLAGS $INC -c "$HERE/ref_preview_capi.cpp" -o "$OUT/obj/ref_preview_capi.o"
$CXX -pthread -shared -o "$OUT/libref_preview.so" "$OUT/obj/ref_preview_capi.o" $OBJS -lz
# the reference's soft ISP: camera_isp/CameraIsp.h + util/JsonUtil.cpp + the JSON library vendored in the reference tree
$CXX $CXXFLAGS $INC -I$REF/camera_isp -c "$HERE/ref_isp_capi.cpp" -o "$OUT/obj/ref_isp_capi.o" &
$CXX $CXXFLAGS $INC -c "$REF/util/JsonUtil.cpp" -o "$OUT/obj/util_JsonUtil.o" &
$CXX $CXXFLAGS $INC -c "$REF/thirdparty/supereasyjson/json.cpp" -o "$OUT/obj/supereasyjson.o" &
wait
$CXX -pthread -shared -o "$OUT/libref_isp.so" "$OUT/obj/ref_isp_capi.o" "$OUT/obj/util_JsonUtil.o" "$OUT/obj/supereasyjson.o" $OBJS -lz
if [ -f "$REPO/surround360_b200/libs360b200.so" ]; then   # the same library with the GPU adapter of integration/reference_tree/camera_isp compiled in
  $CXX $CXXFLAGS $INC -I$REF/camera_isp -I$REPO/integration/reference_tree/camera_isp -I$REPO/include -DREF_WITH_B200_ISP -c "$HERE/ref_isp_capi.cpp" -o "$OUT/obj/ref_isp_capi_b200.o"
  $CXX -pthread -shared -o "$OUT/libref_isp_b200.so" "$OUT/obj/ref_isp_capi_b200.o" "$OUT/obj/util_JsonUtil.o" "$OUT/obj/supereasyjson.o" $OBJS -lz -L"$REPO/surround360_b200" -ls360b200 -Wl,-rpath,'$ORIGIN/../../surround360_b200'
fi
# the Unpacker's front end: the .bin footage container and the 8 / 12-bit sample converter. The two headers use uint32_t /
# CHAR_BIT without including <cstdint> / <climits> (older libstdc++ pulled them in through <string>): pre-included on the command line.
UNP="-include cstdint -include climits -I$REF/camera_isp"
$CXX $CXXFLAGS $INC $UNP -c "$HERE/ref_unpack_capi.cpp" -o "$OUT/obj/ref_unpack_capi.o" &
$CXX $CXXFLAGS $INC $UNP -c "$REF/camera_isp/BinaryFootageFile.cpp" -o "$OUT/obj/camera_isp_BinaryFootageFile.o" &
$CXX $CXXFLAGS $INC $UNP -c "$REF/camera_isp/RawConverter.cpp" -o "$OUT/obj/camera_isp_RawConverter.o" &
wait
$CXX -pthread -shared -o "$OUT/libref_unpack.so" "$OUT/obj/ref_unpack_capi.o" "$OUT/obj/camera_isp_BinaryFootageFile.o" "$OUT/obj/camera_isp_RawConverter.o"
echo "ref_build: built $OUT/TestRenderStereoPanorama and $OUT/libref.so"
