// C-ABI: tode_peer_push -- the dense-output block of this rank's shard goes to every peer's gathered buffer
// by SM stores over NVLink (the all-gather of solutions).
//
// The fused solve kernel writes a sample's statistics to all replicas itself (tode_solution.peer_*), but its
// dense-output rows are 8-byte stores scattered over the sample's integration: over NVLink they are slow
// (DESIGN.md section 8), so the finished block is shipped in bulk afterwards.  Between two GPUs the copy
// engines do that at 750 GB/s; with N - 1 copies in flight in each direction they drop to 300-340 GB/s per
// rank and NCCL's all-gather reaches 510-590 GB/s.  This kernel reads every 16-byte vector of the block once
// from local HBM and stores it to all peer mappings: one pass, every link of the GPU busy at the same time,
// nothing staged, no protocol -- arrival is covered by the cross-GPU barrier the caller runs anyway.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/torchode_b200.h"

namespace {

struct PushArgs {
  const uint4* src;
  uint4* dst[TODE_MAX_PEERS];
  int n_dst;
  long long n_vec;  // 16-byte vectors
};

constexpr int kPushThreads = 256;
constexpr int kPushUnroll = 4;

__global__ void __launch_bounds__(kPushThreads) peer_push_kernel(const __grid_constant__ PushArgs A) {
  const long long stride = (long long)gridDim.x * kPushThreads;
  long long i = (long long)blockIdx.x * kPushThreads + threadIdx.x;
  for (; i + (kPushUnroll - 1) * stride < A.n_vec; i += kPushUnroll * stride) {
    uint4 v[kPushUnroll];
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) v[u] = __ldcs(A.src + i + u * stride);  // read once: streaming
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u)
      for (int p = 0; p < A.n_dst; ++p) A.dst[p][i + u * stride] = v[u];
  }
  for (; i < A.n_vec; i += stride) {
    const uint4 v = __ldcs(A.src + i);
    for (int p = 0; p < A.n_dst; ++p) A.dst[p][i] = v;
  }
  __threadfence_system();
}

}  // namespace

extern "C" int tode_peer_push(const void* src, void* const* dst, int32_t n_dst, int64_t bytes, void* stream) {
  if (!src || !dst || n_dst < 0 || n_dst > TODE_MAX_PEERS || bytes < 0) return TODE_EINVAL;
  if (n_dst == 0 || bytes == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15)) return TODE_EALIGN;
  PushArgs a{};
  a.src = static_cast<const uint4*>(src);
  a.n_dst = n_dst;
  a.n_vec = bytes / 16;
  for (int p = 0; p < n_dst; ++p) {
    if (!dst[p]) return TODE_EINVAL;
    if (reinterpret_cast<uintptr_t>(dst[p]) & 15) return TODE_EALIGN;
    a.dst[p] = static_cast<uint4*>(dst[p]);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long need = (a.n_vec + (long long)kPushThreads * kPushUnroll - 1) / ((long long)kPushThreads * kPushUnroll);
  const long long cap = (long long)sms * 4;
  const unsigned grid = (unsigned)(need < 1 ? 1 : (need > cap ? cap : need));
  peer_push_kernel<<<grid, kPushThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
