// Probe: where do the rows of an M=64 (cta_group::1, kind::f16) UMMA accumulator land in TMEM?
// A[r][k] = (k == 0) ? r + 1 : 0,  B[n][k] = (k == 0) ? 1 : 0  ->  D[r][n] = r + 1 for every n.
// Every warp reads its 32 TMEM lanes (tcgen05.ld.32x32b.x8 at column 0) and prints lane -> value.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_m64_layout umma_m64_layout.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kM = 64, kN = 256, kK = 64;  // one 64-element K-block (4 MMAs of K=16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(128, 1) probe(float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 64 rows x 128 B
  uint8_t* sB = smem + 16384;         // 256 rows x 128 B
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (tid < kM) {
    __nv_bfloat16 v = __float2bfloat16((float)(tid + 1));
    *reinterpret_cast<__nv_bfloat16*>(sA + swz(tid, 0)) = v;  // element k = 0 of row tid
  }
  for (int n = tid; n < kN; n += 128) *reinterpret_cast<__nv_bfloat16*>(sB + swz(n, 0)) = __float2bfloat16(1.0f);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(kM, kN);
    for (int ks = 0; ks < kK / 16; ++ks) {
      const uint64_t a = make_desc(smem_u32(sA) + ks * 32), b = make_desc(smem_u32(sB) + ks * 32);
      const uint32_t acc = ks != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(mbar)), "r"(0u) : "memory");
  } while (!ok);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t r[8];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  out[tid * 2 + 0] = __uint_as_float(r[0]);
  out[tid * 2 + 1] = __uint_as_float(r[7]);
  // second read: columns 128.. (in case M=64 splits the N range across lane halves)
  const uint32_t taddr2 = taddr + 128;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr2) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  out[256 + tid] = __uint_as_float(r[0]);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 384 * sizeof(float));
  cudaMemset(d, 0, 384 * sizeof(float));
  const int smem = 16384 + 32768 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  float h[384];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 4; ++w) {
    printf("warp %d col0  :", w);
    for (int l = 0; l < 32; ++l) printf(" %g", h[(w * 32 + l) * 2]);
    printf("\nwarp %d col7  :", w);
    for (int l = 0; l < 32; ++l) printf(" %g", h[(w * 32 + l) * 2 + 1]);
    printf("\nwarp %d col128:", w);
    for (int l = 0; l < 32; ++l) printf(" %g", h[256 + w * 32 + l]);
    printf("\n");
  }
  return 0;
}
