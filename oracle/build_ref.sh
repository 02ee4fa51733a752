#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference extension
# (facebookresearch/FBTT-Embedding: tt_embeddings.cpp + tt_embeddings_cuda.cu)
# for sm_100a from the sources where they lie under /root/reference, into
# oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
#
# The product never loads this library.  Only tests/, __graft_entry__.smoke()
# and bench.py's reference legs may import it, and only as the checker /
# the "reference CUDA kernels recompiled for sm_100a" baseline.
#
# We do not run the reference's setup.py (it pins compute_70 and a cub-1.8.0
# include path that no longer exists); the three commands below are the same
# translation units with an sm_100a -gencode.  No reference source is copied.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
SO="$OUT/tt_embeddings.cpython-312-x86_64-linux-gnu.so"
if [ ! -d "$REF" ]; then
  echo "[oracle/build_ref] $REF absent (GPU box?) -- using prebuilt $SO if present"; exit 0
fi
# The reference's own Python (module, tests, benchmark) rides along in the same git-ignored directory so that the
# GPU box can run the reference's UNMODIFIED test-suite and benchmark through the `import tt_embeddings` seam
# (tests/test_gpu_reference_seam.py).  Staged copies of files that stay where they lie; never tracked.
mkdir -p "$OUT/py"
for f in tt_embeddings_ops.py tt_embeddings_test.py tt_embeddings_benchmark.py; do
  cp -f "$REF/$f" "$OUT/py/$f"
done
if [ -f "$SO" ] && [ "$SO" -nt "$REF/tt_embeddings_cuda.cu" ] && [ -z "${FORCE:-}" ]; then
  echo "[oracle/build_ref] up to date: $SO"; exit 0
fi
mkdir -p "$OUT"
PY=${PYTHON:-python}
TORCH=$($PY -c "import torch,os;print(os.path.dirname(torch.__file__))")
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
INC="-I$TORCH/include -I$TORCH/include/torch/csrc/api/include -I$PYINC -I/usr/local/cuda/include"
DEFS="-DTORCH_EXTENSION_NAME=tt_embeddings -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1"
(
  nvcc -O3 --expt-relaxed-constexpr -D__CUDA_NO_HALF_OPERATORS__ -std=c++17 -w \
     -gencode=arch=compute_100a,code=sm_100a $INC $DEFS \
     --compiler-options -fPIC -c "$REF/tt_embeddings_cuda.cu" -o "$OUT/tt_embeddings_cuda.o"
) &
(
  g++ -O3 -fPIC -std=c++17 -w $INC $DEFS -c "$REF/tt_embeddings.cpp" -o "$OUT/tt_embeddings.o"
) &
wait
g++ -shared "$OUT/tt_embeddings.o" "$OUT/tt_embeddings_cuda.o" \
    -L"$TORCH/lib" -L/usr/local/cuda/lib64 -Wl,-rpath,"$TORCH/lib" \
    -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -lcudart -lcublas \
    -o "$SO"
rm -f "$OUT"/*.o
echo "[oracle/build_ref] built $SO"
