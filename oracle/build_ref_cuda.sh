#!/usr/bin/env bash
# Compile the reference's registration classes AND its vendored CUDA K-NN (knn.cu, unmodified, from where they lie under
# /root/reference) against the container's libtorch WITH CUDA -- no device swap: this is the reference as it runs on a GPU.
# Output: oracle/_ref/libsvnicp_ref_cuda.so (git-ignored, travels via gpurun).  Used only by bench.py's
# `reference_on_gpu` comparator: the reference itself timed on the same B200.
# TEST/MEASUREMENT INFRASTRUCTURE ONLY.  No reference source is copied into the repo.
set -euo pipefail
REF="${1:-/root/reference/svn-icp}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/core" ]; then echo "reference not present at $REF; keeping prebuilt $OUT" >&2; exit 0; fi
PY="${PYTHON:-python}"
TORCH="$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
ABI="$($PY -c 'import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
mkdir -p "$OUT/cuda"
DEFS="-D_GLIBCXX_USE_CXX11_ABI=$ABI -DREF_CUDA"
INC="-I $HERE/ref_shim -I $REF/include -I $TORCH/include -I $TORCH/include/torch/csrc/api/include -I $PYINC -I $CUDA/include"
CXXFLAGS="-O2 -std=c++17 -fPIC -w $DEFS"
build_cxx() { if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then echo "  CXX $1"; /usr/bin/g++ $CXXFLAGS $INC -c "$1" -o "$2"; fi; }
build_cxx "$REF/src/core/SVGDICP.cpp" "$OUT/cuda/SVGDICP.o" &
build_cxx "$REF/src/core/SVNICP.cpp" "$OUT/cuda/SVNICP.o" &
build_cxx "$REF/src/core/knn/knn.cpp" "$OUT/cuda/knn.o" &
build_cxx "$REF/src/core/knn/knn_cpu.cpp" "$OUT/cuda/knn_cpu.o" &
build_cxx "$HERE/ref_driver.cpp" "$OUT/cuda/ref_driver.o" &
if [ ! -f "$OUT/cuda/knn_cu.o" ] || [ "$REF/src/core/knn/knn.cu" -nt "$OUT/cuda/knn_cu.o" ]; then
  echo "  NVCC $REF/src/core/knn/knn.cu"
  "$CUDA/bin/nvcc" -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC -w $DEFS $INC \
      --expt-relaxed-constexpr -c "$REF/src/core/knn/knn.cu" -o "$OUT/cuda/knn_cu.o" &
fi
wait
/usr/bin/g++ -shared -o "$OUT/libsvnicp_ref_cuda.so" "$OUT/cuda/SVGDICP.o" "$OUT/cuda/SVNICP.o" "$OUT/cuda/knn.o" "$OUT/cuda/knn_cpu.o" \
    "$OUT/cuda/knn_cu.o" "$OUT/cuda/ref_driver.o" -L "$TORCH/lib" -Wl,-rpath,"$TORCH/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda \
    -ltorch_python -L "$CUDA/lib64" -lcudart
echo "built $OUT/libsvnicp_ref_cuda.so"
