// ORACLE build glue (test infrastructure, NOT product code): exposes the REFERENCE's own
// CUTLASS GEGLU kernel -- /root/reference/src/sfast/csrc/operators/cutlass/
// cutlass_dual_linear_kernel.cu, compiled from where it lies by oracle/build_ref.sh -- as
// torch.ops.sfast_ref.linear_geglu(input, weight, bias) so that tests/ can run it on the B200.
//
// The reference registers the same function as torch.ops.sfast.cutlass_linear_geglu_unified
// (/root/reference/src/sfast/csrc/operators/cutlass/cutlass_dual_linear.cc); its shape/dtype
// fallback calls cublas_lowp_linear, whose translation unit (operators/cublas/CUDABlas.cc) does
// not compile against torch 2.11, so that one symbol is provided here as
// plain at::linear -- it is only reached for dtypes / alignments the CUTLASS kernel refuses.
#include <torch/extension.h>
#include <torch/library.h>

#include "cutlass_dual_linear_kernel.h"
#include "operators/cublas/cublas_gemm.h"

namespace sfast {
namespace operators {
torch::Tensor cublas_lowp_linear(const torch::Tensor &input, const torch::Tensor &weight,
                                 const c10::optional<torch::Tensor> &bias) {
  return at::linear(input, weight, bias);
}
}  // namespace operators
}  // namespace sfast

TORCH_LIBRARY(sfast_ref, m) {
  m.def("linear_geglu(Tensor input, Tensor weight, Tensor? bias) -> Tensor",
        [](const torch::Tensor &input, const torch::Tensor &weight,
           const c10::optional<torch::Tensor> &bias) {
          return sfast::operators::cutlass_linear_geglu_unified(input, weight, bias);
        });
}
