#!/usr/bin/env bash
# Compile the reference's own registration sources (unmodified, from where they lie under
# /root/reference) against the container's libtorch, device-swapped to CPU, together with
# oracle/ref_driver.cpp.  Output: oracle/_ref/libsvnicp_ref.so (git-ignored, travels via gpurun).
# Recipe: SURVEY.md Appendix B.  No reference source is copied into the repo.
set -euo pipefail
REF="${1:-/root/reference/svn-icp}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/core" ]; then echo "reference not present at $REF; keeping prebuilt $OUT" >&2; exit 0; fi
PY="${PYTHON:-python}"
TORCH="$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
mkdir -p "$OUT"
CXXFLAGS="-O2 -std=c++17 -fPIC -DSVN_ORACLE_CPU -D_GLIBCXX_USE_CXX11_ABI=$($PY -c 'import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))') -w"
INC="-I $HERE/ref_shim -I $REF/include -I $TORCH/include -I $TORCH/include/torch/csrc/api/include -I $PYINC"
build_one() { # src obj
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then echo "  CXX $1"; g++ $CXXFLAGS $INC -c "$1" -o "$2"; fi
}
build_one "$REF/src/core/SVGDICP.cpp" "$OUT/SVGDICP.o" &
build_one "$REF/src/core/SVNICP.cpp" "$OUT/SVNICP.o" &
build_one "$HERE/ref_driver.cpp" "$OUT/ref_driver.o" &
wait
g++ -shared -o "$OUT/libsvnicp_ref.so" "$OUT/SVGDICP.o" "$OUT/SVNICP.o" "$OUT/ref_driver.o" \
    -L "$TORCH/lib" -Wl,-rpath,"$TORCH/lib" -ltorch -ltorch_cpu -lc10
echo "built $OUT/libsvnicp_ref.so"
# the reference's local map (VoxelHashMap.cpp, unmodified) over the stand-in PCL / Eigen / tsl / gtsam types of ref_shim_map
g++ -O2 -std=c++17 -fPIC -w -I "$HERE/ref_shim_map" -I "$REF/include" -shared -o "$OUT/libvmap_ref.so" \
    "$REF/src/core/VoxelHashMap.cpp" "$HERE/ref_driver_map.cpp"
echo "built $OUT/libvmap_ref.so"
