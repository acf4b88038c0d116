// The stencil of fields.Heat1D, shared by the stand-alone field kernel (heat_field.cu) and the
// stage-fused step kernel (heat_step.cu): same operation order, one IEEE rounding per operation, as
//   kappa * ((y[:, 2:] - 2 * y[:, 1:-1]) + y[:, :-2])
#pragma once

namespace tode {
namespace heat {

__device__ __forceinline__ float stencil(float l, float c, float r, float kappa) {
  return __fmul_rn(kappa, __fadd_rn(__fsub_rn(r, __fmul_rn(2.0f, c)), l));
}
__device__ __forceinline__ double stencil(double l, double c, double r, double kappa) {
  return __dmul_rn(kappa, __dadd_rn(__dsub_rn(r, __dmul_rn(2.0, c)), l));
}

}  // namespace heat
}  // namespace tode
