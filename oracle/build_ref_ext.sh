#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY — builds the reference's own CUDA extensions as a parity oracle.
#
# Compiles the four reference extension source pairs (src/<name>.cu + src/bindings.cpp) IN PLACE from
# /root/reference (nothing is copied into this repo) for sm_100a, writing only into oracle/_ref/
# (git-ignored, but shipped to the GPU box by gpurun).  The single deviation from the reference's own
# flags (gridencoder/backend.py:6-12) is -std=c++17 instead of -std=c++14, which torch 2.11 headers
# require, plus --expt-relaxed-constexpr which torch.utils.cpp_extension always passes.
# The resulting modules are imported ONLY by tests/ (GPU parity) and bench.py's reference-ext timing leg.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF" ]; then echo "[build_ref_ext] $REF absent; using prebuilt $OUT if any"; exit 0; fi
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))' 2>/dev/null)
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
EXT=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
INC="-I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC -I/usr/local/cuda/include"
DEFS="-DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1"
NVF="-O3 -std=c++17 --expt-relaxed-constexpr -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__ -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
build_one() {  # $1 = package dir, $2 = .cu stem, $3 = module name, $4 = extra nvcc flags
  local pkg=$1 stem=$2 mod=$3 extra=${4:-}
  local so="$OUT/$mod$EXT"
  if [ -f "$so" ] && [ "$so" -nt "$REF/$pkg/src/$stem.cu" ]; then echo "[build_ref_ext] $mod up to date"; return 0; fi
  local tmp; tmp=$(mktemp -d)
  nvcc $NVF $extra $INC $DEFS -DTORCH_EXTENSION_NAME=$mod -c "$REF/$pkg/src/$stem.cu" -o "$tmp/$stem.o"
  g++ -O3 -std=c++17 -fPIC $INC $DEFS -DTORCH_EXTENSION_NAME=$mod -c "$REF/$pkg/src/bindings.cpp" -o "$tmp/bindings.o"
  g++ -shared "$tmp/$stem.o" "$tmp/bindings.o" -L"$TORCH_DIR/lib" -L/usr/local/cuda/lib64 \
      -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -lcudart \
      -Wl,-rpath,"$TORCH_DIR/lib" -o "$so"
  rm -rf "$tmp"
  echo "[build_ref_ext] built $so"
}
# module names are the reference's JIT names (*/backend.py:31)
build_one raymarching raymarching _ref_raymarching_face &
build_one gridencoder gridencoder _ref_grid_encoder &
build_one shencoder   shencoder   _ref_sh_encoder &
build_one freqencoder freqencoder _ref_freqencoder "-use_fast_math" &
wait
ls -la "$OUT"
