// C-ABI: version, error strings, scratch sizing.
#include <cuda_runtime.h>

#include "../../include/torchode_b200.h"

extern "C" int tode_abi_version(void) { return TODE_ABI_VERSION; }

extern "C" const char* tode_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case TODE_EINVAL: return "invalid argument (NULL pointer, bad size or enum)";
    case TODE_ENOSUP: return "combination not supported by this build";
    case TODE_EALIGN: return "operand not aligned as the layout contract demands (16 bytes)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

// scratch (data dtype elements): 2*B for the initial step (dt0, d1) + in split mode two partials
// per (sample, 1024-vector chunk); split-mode finish: one partial per (sample, chunk) + a 32-byte
// step record per sample (sized for the narrowest vector / element width so that it holds for
// every dtype)
extern "C" int64_t tode_scratch_elems(int64_t B, int64_t F) {
  const int64_t chunks = (F + 1023) / 1024;
  return 2 * B + 2 * B * chunks + 8 * B + 64;
}

// ---- measurement aid: peak double-precision FMA issue rate (bench.py roofline) ------------
namespace {
constexpr int kPeakBlock = 256;
constexpr int kPeakChains = 8;
__global__ void __launch_bounds__(kPeakBlock) fp64_fma_kernel(long long iters, double* sink) {
  double a[kPeakChains];
  const double x = 1.0000001 + 1e-9 * threadIdx.x, y = 1e-7;
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) a[c] = 1.0 + c;
  for (long long i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) a[c] = __fma_rn(a[c], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) s += a[c];
  sink[(long long)blockIdx.x * kPeakBlock + threadIdx.x] = s;
}
int peak_grid() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n * 8;
}
}  // namespace

extern "C" int64_t tode_bench_fp64_fma_threads(void) { return (int64_t)peak_grid() * kPeakBlock; }

extern "C" int tode_bench_fp64_fma(int64_t iters, void* sink, int64_t* n_fma_out, void* stream) {
  if (!sink || iters <= 0) return TODE_EINVAL;
  const int grid = peak_grid();
  fp64_fma_kernel<<<grid, kPeakBlock, 0, static_cast<cudaStream_t>(stream)>>>(iters, static_cast<double*>(sink));
  if (n_fma_out) *n_fma_out = (int64_t)grid * kPeakBlock * kPeakChains * iters;
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
