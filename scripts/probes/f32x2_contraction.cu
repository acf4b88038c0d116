// Probe (compile-only): does the toolchain keep a packed fp32 multiplication and a following packed addition
// apart?  __fmul2_rn / __fadd2_rn lower to mul.rn.f32x2 / add.rn.f32x2 in PTX; for SCALAR fp32 the .rn forms
// are never contracted.  With nvcc 12.9 -fmad=false, ptxas nevertheless emits ONE FFMA2 for the packed pair
// (one rounding instead of two), while the scalar pair stays FMUL + FADD.  heat_step.cu therefore keeps every
// multiplication whose result feeds an addition scalar (the error estimate) and writes 2*c as c + c.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o /tmp/p.cubin scripts/probes/f32x2_contraction.cu
//   cuobjdump -sass /tmp/p.cubin | grep -E "FFMA2|FMUL2|FADD2|FFMA |FMUL |FADD "
//
// Captured output: profiles/r01_ptxas_f32x2_contraction.txt
#include <cuda_runtime.h>

extern "C" __global__ void packed_mul_then_add(const float2* a, const float2* b, const float2* c, float2* out) {
  const int i = threadIdx.x;
  out[i] = __fadd2_rn(c[i], __fmul2_rn(a[i], b[i]));  // two roundings requested
}

extern "C" __global__ void scalar_mul_then_add(const float* a, const float* b, const float* c, float* out) {
  const int i = threadIdx.x;
  out[i] = __fadd_rn(c[i], __fmul_rn(a[i], b[i]));  // two roundings requested
}

extern "C" __global__ void packed_mul_by_two_then_sub(const float2* r, const float2* c, float2* out) {
  const int i = threadIdx.x;
  const float2 two = make_float2(2.0f, 2.0f);
  const float2 p = __fmul2_rn(two, c[i]);
  out[i] = __fadd2_rn(r[i], make_float2(-p.x, -p.y));  // r - 2c with an overflowing product -> -inf
}

extern "C" __global__ void packed_mul_by_minus_two_then_add(const float2* r, const float2* c, float2* out) {
  const int i = threadIdx.x;
  const float2 m2 = make_float2(-2.0f, -2.0f);
  out[i] = __fadd2_rn(r[i], __fmul2_rn(m2, c[i]));  // the form heat_step.cu used first: contracted as well
}

extern "C" __global__ void packed_double_by_add_then_sub(const float2* r, const float2* c, float2* out) {
  const int i = threadIdx.x;
  const float2 p = __fadd2_rn(c[i], c[i]);
  out[i] = __fadd2_rn(r[i], make_float2(-p.x, -p.y));  // what heat_step.cu does instead
}
