// Method-of-lines vector field of the 1-D heat equation (BASELINE.json configs[4]):
//   out[b, i] = kappa * ((y[b, i+1] - 2 y[b, i]) + y[b, i-1])   for 0 < i < N-1,   out[b, 0] = out[b, N-1] = 0
// (Dirichlet ends).  One HBM pass: every thread owns 16 bytes of a row, the two halo values come
// from its neighbours' cache lines.  Same operation order, one IEEE rounding per operation, as the
// PyTorch expression  kappa * ((y[:, 2:] - 2 * y[:, 1:-1]) + y[:, :-2])  -- a user-level f like the
// MLP field (terms.py:60-63 calls it), not part of the solver arithmetic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/torchode_b200.h"
#include "heat_stencil.cuh"

namespace tode {
namespace heat {

constexpr int kThreads = 256;


// VEC elements (16 bytes) per thread; N % VEC == 0 so a vector never straddles rows
template <typename D, int VEC>
__global__ void __launch_bounds__(kThreads) heat1d_kernel(const D* __restrict__ y, D* __restrict__ out, long long n_vec,
                                                          long long N, D kappa) {
  const long long vec_per_row = N / VEC;
  for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < n_vec; v += (long long)gridDim.x * kThreads) {
    const long long i0 = (v % vec_per_row) * VEC;  // first column of this vector
    const D* p = y + v * VEC;
    D c[VEC + 2];
    if (VEC == 4) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      c[1] = (D)q.x; c[2] = (D)q.y; c[3] = (D)q.z; c[4] = (D)q.w;
    } else {
      const double2 q = *reinterpret_cast<const double2*>(p);
      c[1] = (D)q.x; c[2] = (D)q.y;
    }
    c[0] = i0 > 0 ? p[-1] : (D)0;
    c[VEC + 1] = i0 + VEC < N ? p[VEC] : (D)0;
    D r[VEC];
#pragma unroll
    for (int x = 0; x < VEC; ++x) {
      const long long i = i0 + x;
      r[x] = (i == 0 || i == N - 1) ? (D)0 : stencil(c[x], c[x + 1], c[x + 2], kappa);
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(out + v * VEC) = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
    } else {
      *reinterpret_cast<double2*>(out + v * VEC) = make_double2((double)r[0], (double)r[1]);
    }
  }
}

}  // namespace heat
}  // namespace tode

extern "C" int tode_heat1d_forward(const void* y, void* out, int64_t B, int64_t N, double kappa, int32_t dtype,
                                   void* stream) {
  using namespace tode::heat;
  if (!y || !out || B < 0 || N < 2 || (dtype != TODE_F32 && dtype != TODE_F64)) return TODE_EINVAL;
  if (B == 0) return 0;
  const int vec = dtype == TODE_F32 ? 4 : 2;
  if (N % vec != 0) return TODE_ENOSUP;
  if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return TODE_EALIGN;
  const long long n_vec = B * N / vec;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long need = (n_vec + kThreads - 1) / kThreads;
  const long long cap = (long long)sms * 16;
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == TODE_F32)
    heat1d_kernel<float, 4><<<grid, kThreads, 0, st>>>(static_cast<const float*>(y), static_cast<float*>(out), n_vec, N,
                                                       (float)kappa);
  else
    heat1d_kernel<double, 2><<<grid, kThreads, 0, st>>>(static_cast<const double*>(y), static_cast<double*>(out), n_vec,
                                                        N, kappa);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
