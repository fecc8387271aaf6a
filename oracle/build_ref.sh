#!/bin/bash
# ORACLE build recipe (test infrastructure): compiles the reference's CUTLASS GEGLU kernel from
# its own source file under /root/reference into oracle/_ref/libsfast_ref_geglu.so (git-ignored;
# travels to the GPU box with the snapshot).  Nothing from /root/reference is copied into the repo.
#
# torch 2.11 changed at::Context::allowFP16ReductionCuBLAS()/allowBF16ReductionCuBLAS() from
# `bool` to the enum at::CuBLASReductionOption, which the reference uses as
# a bool at cutlass_dual_linear_kernel.cu:489,509.  /root/reference is read-only, so a scratch copy
# of that ONE file gets the two calls replaced by `true` -- torch's default flag value, i.e. the
# reference's default code path: fp16/bf16 accumulation + GELU_taylor_fast (tanh form).
set -euo pipefail
REF=/root/reference
[ -d "$REF" ] || { echo "no $REF here (GPU box): using the prebuilt oracle/_ref"; exit 0; }
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"; mkdir -p "$OUT"
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
SRC=$REF/src/sfast/csrc/operators/cutlass/cutlass_dual_linear_kernel.cu
sed -e 's/at::globalContext().allowFP16ReductionCuBLAS()/true/' \
    -e 's/at::globalContext().allowBF16ReductionCuBLAS()/true/' "$SRC" > "$TMP/cutlass_dual_linear_kernel.cu"
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_LIB PY_INC CXX11 < <($PY - <<'P'
import sysconfig, torch
from torch.utils import cpp_extension as c
print(" ".join("-I" + p for p in c.include_paths()).replace(" ", ","), c.library_paths()[0],
      sysconfig.get_paths()["include"], int(torch._C._GLIBCXX_USE_CXX11_ABI))
P
)
INC="${TORCH_INC//,/ } -I$PY_INC -I$REF/src/sfast/csrc -I$REF/src/sfast/csrc/operators/cutlass \
 -I$REF/third_party/cutlass/include -I$REF/third_party/cutlass/examples/45_dual_gemm \
 -I$REF/third_party/cutlass/tools/util/include"
FLAGS="-O2 -std=c++17 -Xcompiler -fPIC -DWITH_CUDA -D_GLIBCXX_USE_CXX11_ABI=$CXX11 -DTORCH_EXTENSION_NAME=sfast_ref \
 --expt-relaxed-constexpr --expt-extended-lambda -gencode arch=compute_100a,code=sm_100a -w"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC $FLAGS $INC -c "$TMP/cutlass_dual_linear_kernel.cu" -o "$TMP/kernel.o"
$NVCC $FLAGS $INC -x cu -c "$HERE/ref_geglu_binding.cpp" -o "$TMP/binding.o"
$NVCC -shared -o "$OUT/libsfast_ref_geglu.so" "$TMP/kernel.o" "$TMP/binding.o" \
  -L"$TORCH_LIB" -lc10 -ltorch -ltorch_cpu -ltorch_cuda -lc10_cuda
echo "built $OUT/libsfast_ref_geglu.so"
