// Neural-ODE vector field on the 5th-gen tensor cores: y -> W_L tanh(... tanh(W_1 y + b_1) ...) + b_L
// for a width-256 tanh MLP (BASELINE.json configs[3]).  This is the one place on the solve path
// where the work really is a dense GEMM: bf16 operands, fp32 accumulation in TMEM (tcgen05.mma
// issued by one elected thread), bias + tanh epilogue out of TMEM (tcgen05.ld), the activation
// tile of a CTA never leaves shared memory between layers.
//
// One CTA = 128 rows of the batch (UMMA M = 128, N = 256, K = 16 per instruction, 16 per layer).
// Shared memory holds the activation tile (128 x 256 bf16, 64 KB) and one layer's weights
// (256 x 256 bf16, 128 KB), both K-major in the canonical SWIZZLE_128B layout: K-blocks of 64
// elements (128 B rows), 8-row atoms of 1024 B (SBO), 16-byte chunks XOR-swizzled with row % 8.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/torchode_b200.h"

namespace tode {
namespace mlp {

constexpr int kWidth = 256;
constexpr int kBM = 128;                       // rows per CTA (UMMA M)
constexpr int kThreads = 512;                  // 16 warps: 4 TMEM lane quarters x 4 column quarters
constexpr int kKBlock = 64;                    // bf16 elements per 128-byte swizzle row
constexpr int kNumKBlocks = kWidth / kKBlock;  // 4
constexpr int kABlockBytes = kBM * 128;        // 16 KB per K-block of the activation tile
constexpr int kWBlockBytes = kWidth * 128;     // 32 KB per K-block of the weight tile
constexpr int kSmemA = kNumKBlocks * kABlockBytes;  // 64 KB
constexpr int kSmemW = kNumKBlocks * kWBlockBytes;  // 128 KB
constexpr int kSmemBytes = kSmemA + kSmemW + kWidth * 4 + 64;  // + bias + barrier / TMEM slot
constexpr uint32_t kTmemCols = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 UMMA): start address >> 4 in
// bits [0,14), stride byte offset (8 rows x 128 B = 1024) >> 4 in bits [32,46), descriptor
// version 1 in bits [46,48), layout type SWIZZLE_128B = 2 in bits [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// byte offset of the 16-byte chunk holding elements [8c, 8c+8) of K-block kb of `row`
__device__ __forceinline__ uint32_t swz(uint32_t block_bytes, int kb, int row, int c) {
  return (uint32_t)kb * block_bytes + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
}

// NB: weights are stored [layer][out][in] = (N, K) row-major, i.e. already K-major for the B operand
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
mlp_tanh256_kernel(const float* __restrict__ y, const __nv_bfloat16* __restrict__ weights,
                   const float* __restrict__ biases, float* __restrict__ out, long long B, int n_layers) {
  // 1024-byte alignment (SWIZZLE_128B atoms) is requested from the compiler / driver; no integer
  // round-up, so that the compiler keeps these pointers in the shared address space (LDS/STS)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kSmemA;
  float* sBias = reinterpret_cast<float*>(smem + kSmemA + kSmemW);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kSmemA + kSmemW + kWidth * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * kBM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  // ---- layer 0 weights: asynchronous global -> shared copies, in flight during the A load ----
  auto load_weights_async = [&](int layer) {
    const uint4* wsrc = reinterpret_cast<const uint4*>(weights + (size_t)layer * kWidth * kWidth);
    const uint32_t sW_base = smem_u32(sW);
#pragma unroll 8
    for (int idx = tid; idx < kWidth * (kWidth / 8); idx += kThreads) {
      const int n = idx / (kWidth / 8), chunk = idx % (kWidth / 8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sW_base + swz(kWBlockBytes, chunk >> 3, n, chunk & 7)),
                   "l"(wsrc + idx)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  load_weights_async(0);

  // ---- activation tile: fp32 rows of y -> bf16, swizzled K-major (rows past B are zero) ----
  constexpr int kAChunks = kBM * (kWidth / 8);  // 4096 chunks of 8 elements, 16 per thread
  constexpr int kAU = 8;                        // chunks in flight per thread (16 x 16-byte loads)
#pragma unroll 1
  for (int base = 0; base < kAChunks; base += kThreads * kAU) {
    float4 v0[kAU], v1[kAU];
#pragma unroll
    for (int u = 0; u < kAU; ++u) {  // all loads first (memory-level parallelism)
      const int idx = base + u * kThreads + tid;
      const int row = idx / (kWidth / 8), chunk = idx % (kWidth / 8);
      v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      v1[u] = v0[u];
      if (m0 + row < B) {
        const float4* src = reinterpret_cast<const float4*>(y + (m0 + row) * kWidth + chunk * 8);
        v0[u] = src[0];
        v1[u] = src[1];
      }
    }
#pragma unroll
    for (int u = 0; u < kAU; ++u) {
      const int idx = base + u * kThreads + tid;
      const int row = idx / (kWidth / 8), chunk = idx % (kWidth / 8);
      uint4 p;
      p.x = pack_bf16(v0[u].x, v0[u].y);
      p.y = pack_bf16(v0[u].z, v0[u].w);
      p.z = pack_bf16(v1[u].x, v1[u].y);
      p.w = pack_bf16(v1[u].z, v1[u].w);
      *reinterpret_cast<uint4*>(sA + swz(kABlockBytes, chunk >> 3, row, chunk & 7)) = p;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(kBM, kWidth);
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW), bar = smem_u32(mbar);
  uint32_t parity = 0;

  for (int layer = 0; layer < n_layers; ++layer) {
    // ---- this layer's weights (out, in) = (N, K) row-major -> K-major swizzled, and bias ------
    // (issued right after the previous layer's MMAs completed, in flight during its epilogue)
    if (tid < kWidth) sBias[tid] = biases[layer * kWidth + tid];
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    // generic-proxy smem writes (cp.async, st.shared) -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();

    if (warp == 0 && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int kb = 0; kb < kNumKBlocks; ++kb) {
#pragma unroll
        for (int ks = 0; ks < kKBlock / 16; ++ks) {
          const uint64_t a_desc = make_desc(sA_addr + kb * kABlockBytes + ks * 32);
          const uint64_t b_desc = make_desc(sW_addr + kb * kWBlockBytes + ks * 32);
          mma_bf16(tmem_base, a_desc, b_desc, idesc, (kb | ks) != 0 ? 1u : 0u);
        }
      }
      // arrives on the mbarrier when every MMA above has completed (implies fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
                   : "memory");
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // the tensor core is done reading sW: fetch the next layer's weights behind the epilogue
    if (layer + 1 < n_layers) load_weights_async(layer + 1);

    // ---- epilogue: TMEM -> registers, + bias, tanh, -> next layer's activation tile / out ----
    const int q = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter
    const int row = q * 32 + lane;
    const bool last = layer == n_layers - 1;
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      const int n0 = cq * 64 + j * 32;
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
            "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
            "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r[i]) + sBias[n0 + i];
        // hidden activations are rounded to bf16 (2^-9) right after: the hardware tanh
        // approximation (max rel. error 2^-11) is below that resolution
        if (!last) asm("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      }
      if (!last) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 8 columns = one 16-byte chunk of the next layer's K
          const int n = n0 + g * 8;
          uint4 p;
          p.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
          p.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
          p.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
          p.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
          *reinterpret_cast<uint4*>(sA + swz(kABlockBytes, n >> 6, row, (n & 63) >> 3)) = p;
        }
      } else {
        // last layer: stage the fp32 tile in the (now idle) weight buffer, XOR-swizzled by row so
        // that neither these per-row writes nor the per-column reads below conflict on a bank
        float* sOut = reinterpret_cast<float*>(sW);
#pragma unroll
        for (int i = 0; i < 32; ++i) sOut[row * kWidth + ((n0 + i) ^ (row & 31))] = v[i];
      }
    }
    if (last) {
      __syncthreads();
      // coalesced copy-out: one warp writes whole 1 KB rows, 128 contiguous bytes per instruction
      const float* sOut = reinterpret_cast<const float*>(sW);
      for (int r = warp; r < kBM; r += kThreads / 32) {
        if (m0 + r < B) {
          float* dst = out + (m0 + r) * kWidth;
#pragma unroll
          for (int j = 0; j < kWidth / 32; ++j) dst[lane + 32 * j] = sOut[r * kWidth + ((lane + 32 * j) ^ (r & 31))];
        }
      }
    }
    // TMEM reads and smem writes of this layer are done before the next layer's MMA starts
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
  }

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace mlp
}  // namespace tode

extern "C" int tode_mlp_tanh256_forward(const void* y, const void* weights_bf16, const void* biases_f32,
                                        void* out, int64_t B, int32_t n_layers, void* stream) {
  using namespace tode::mlp;
  if (!y || !weights_bf16 || !biases_f32 || !out || n_layers < 1 || n_layers > 8) return TODE_EINVAL;
  if (B == 0) return 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(y) || !al16(weights_bf16) || !al16(out)) return TODE_EALIGN;
  static bool configured = false;
  if (!configured) {
    const cudaError_t e = cudaFuncSetAttribute(mlp_tanh256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const unsigned grid = (unsigned)((B + kBM - 1) / kBM);
  mlp_tanh256_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(y), static_cast<const __nv_bfloat16*>(weights_bf16),
      static_cast<const float*>(biases_f32), static_cast<float*>(out), (long long)B, (int)n_layers);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
