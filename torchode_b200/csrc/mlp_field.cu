// Neural-ODE vector field on the 5th-gen tensor cores: y -> W_L tanh(... tanh(W_1 y + b_1) ...) + b_L
// for a width-256 tanh MLP (BASELINE.json configs[3]).  This is the one place on the solve path
// where the work really is a dense GEMM: bf16 operands, fp32 accumulation in TMEM (tcgen05.mma
// issued by one elected thread), bias + tanh epilogue out of TMEM (tcgen05.ld), the activation
// tile of a CTA never leaves shared memory between layers.
//
// One CTA = BM rows of the batch (UMMA M = BM, N = 256, K = 16 per instruction, 16 per layer).
// Shared memory holds the activation tile (BM x 256 bf16) and one layer's weights (256 x 256
// bf16, 128 KB), both K-major in the canonical SWIZZLE_128B layout: K-blocks of 64 elements
// (128 B rows), 8-row atoms of 1024 B (SBO), 16-byte chunks XOR-swizzled with row % 8.
//
// BM = 128 for large batches; BM = 64 when 128-row tiles would leave SMs idle (B = 8192 of
// configs[3]: 64 CTAs on 148 SMs).  An M = 64 UMMA costs the tensor core as much as M = 128, but the
// epilogue -- bound by MUFU.TANH, 16 lanes / clk / SM, 51 % of the kernel at BM = 128 (phase
// stamps of -DTODE_MLP_TIMING) -- and the tile loads halve per CTA while twice as many SMs work.
// M = 64 accumulators occupy lanes 0..15 of each TMEM subpartition (row r -> lane 32 (r / 16) +
// r % 16, probed with scripts/probes/umma_m64_layout.cu); they are read with tcgen05.ld.16x256b,
// which spreads 16 lanes over all 32 threads in the mma C-fragment layout (thread l, repeat j:
// rows l / 4 and l / 4 + 8, columns 8 j + 2 (l % 4) + {0, 1}; scripts/probes/umma_m64_ld16x256b.cu).
#include <cuda.h>  // CUtensorMap (the driver entry point is resolved at run time: no link against libcuda)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/torchode_b200.h"

// Weight tiles travel by TMA (cp.async.bulk.tensor, one thread issues two 32 KB boxes per half; the hardware writes
// the SWIZZLE_128B layout the UMMA descriptors expect and signals an mbarrier; round 1 used 512 threads x 8 cp.async)
#ifndef TODE_MLP_PDL_DEFAULT
#define TODE_MLP_PDL_DEFAULT 0
#endif

namespace tode {
namespace mlp {

constexpr int kWidth = 256;
constexpr int kThreads = 512;                  // 16 warps: 4 TMEM lane quarters x 4 column quarters
constexpr int kKBlock = 64;                    // bf16 elements per 128-byte swizzle row
constexpr int kNumKBlocks = kWidth / kKBlock;  // 4
constexpr int kWBlockBytes = kWidth * 128;     // 32 KB per K-block of the weight tile
constexpr int kSmemW = kNumKBlocks * kWBlockBytes;  // 128 KB
// TMEM columns: the accumulator (256); step-fused launches keep the fp32 y tile of the step in 128 more
__host__ __device__ constexpr uint32_t tmem_cols(bool step) { return step ? 512u : 256u; }
// BM = rows per CTA (UMMA M): the activation tile takes BM * 128 B per K-block
__host__ __device__ constexpr int smem_a_bytes(int bm) { return kNumKBlocks * bm * 128; }
constexpr int kMaxLayers = 8;
// behind the two tiles: BM = 128 keeps every layer's bias (8 KB); BM = 64 keeps the fp32 partial stage sum of the
// NEXT stage's operand rows (64 x 256 fp32 = 64 KB, step-fused launches) and reads its biases into registers
constexpr int kBiasBytes = kMaxLayers * kWidth * 4;
constexpr int kPartialBytes = 64 * kWidth * 4;
__host__ __device__ constexpr int smem_extra_bytes(int bm) { return bm == 64 ? kPartialBytes : kBiasBytes; }
__host__ __device__ constexpr int smem_bytes(int bm) { return smem_a_bytes(bm) + kSmemW + smem_extra_bytes(bm) + 64; }  // + 3 barriers + TMEM slot

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

// one box of a tiled tensor map -> shared memory; completion is counted in bytes on `mbar`
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t mbar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_dst),
      "l"(tmap), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
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

// 32 consecutive columns of this thread's TMEM lane <-> 32 registers
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
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
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
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

// physical column of logical column `col` of row `row` in the staged fp32 output tile: a
// per-row permutation that makes both the epilogue's writes and the row-wise copy-out
// conflict-free (BM = 128: one row per lane -> XOR with the row; BM = 64: the C-fragment layout
// has 8 rows x 4 column pairs per instruction: 8-byte stores, a half-warp = 4 rows x 4 pairs -> rotation by
// 8 (row % 4) columns, which keeps 4-column groups together for the 16-byte reads of the copy-out)
template <int BM>
__device__ __forceinline__ int out_col(int row, int col) {
  if (BM == 128) return col ^ (row & 31);
  return (col + 8 * (row & 3)) & (kWidth - 1);
}

__device__ __forceinline__ float bias_act(uint32_t acc, float bias, bool last) {
  float v = __uint_as_float(acc) + bias;
  // hidden activations are rounded to bf16 (2^-9) right after: the hardware tanh
  // approximation (max rel. error 2^-11) is below that resolution
  if (!last) asm("tanh.approx.f32 %0, %0;" : "+f"(v));
  return v;
}

// BM = 128: row = TMEM lane; every warp reads its 32 lanes x 64 columns in two 32-column pieces
__device__ __forceinline__ void epilogue_m128(uint32_t tmem_base, uint8_t* sA, uint8_t* sW, const float* sBias,
                                              int q, int cq, int lane, bool last) {
  constexpr int kABlockBytes = 128 * 128;
  const int row = q * 32 + lane;
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
    for (int i = 0; i < 32; ++i) v[i] = bias_act(r[i], sBias[n0 + i], last);
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
      // last layer: stage the fp32 tile in the (now idle) weight buffer
      float* sOut = reinterpret_cast<float*>(sW);
#pragma unroll
      for (int i = 0; i < 32; ++i) sOut[row * kWidth + out_col<128>(row, n0 + i)] = v[i];
    }
  }
}

// BM = 64: rows 16 q .. 16 q + 15 sit in lanes 0..15 of subpartition q; 16x256b.x8 hands thread
// `lane` the rows lane / 4 (+ 8) and, per repeat j, the column pair 8 j + 2 (lane % 4) + {0, 1}
// of this warp's 64 columns: 32 elements per thread, all 32 threads busy
// `bia[j]`: the bias of this thread's column pair of repeat j (loaded from global memory before the MMA wait)
__device__ __forceinline__ void epilogue_m64(uint32_t tmem_base, uint8_t* sA, uint8_t* sW, const float2* bia,
                                             int q, int cq, int lane, bool last) {
  constexpr int kABlockBytes = 64 * 128;
  uint32_t r[32];
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
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
  const int row_lo = q * 16 + (lane >> 2), cpair = 2 * (lane & 3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = cq * 64 + 8 * j + cpair;  // and col + 1
    const float b0 = bia[j].x, b1 = bia[j].y;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = row_lo + 8 * h;
      const float v0 = bias_act(r[4 * j + 2 * h], b0, last), v1 = bias_act(r[4 * j + 2 * h + 1], b1, last);
      if (!last) {
        // K-block cq, 16-byte chunk j of the next layer's activation row, 4-byte slot lane % 4
        *reinterpret_cast<uint32_t*>(sA + swz(kABlockBytes, cq, row, j) + 2 * cpair) = pack_bf16(v0, v1);
      } else {
        float* sOut = reinterpret_cast<float*>(sW);
        *reinterpret_cast<float2*>(sOut + row * kWidth + out_col<64>(row, col)) = make_float2(v0, v1);
      }
    }
  }
}

// Stage-fused evaluation (tode_mlp_tanh256_stage_forward): the rows handed to the MLP are not read
// from memory but formed while the activation tile is loaded,
//   y_i = y + dt * sum_{j < nk} a[j] * k[j]        (runge_kutta.py:259-263)
// with the arithmetic of erk_stage_kernel (product, FMA chain in ascending j, one FMA with dt), so
// that the launch replaces tode_erk_stage + tode_mlp_tanh256_forward bit for bit.  nk == 0: plain
// evaluation of y.
//
// Step-fused evaluation (tode_mlp_tanh256_step_forward, round 2): stages stage0 .. stage1 in ONE launch.  f acts
// row by row, so a CTA's rows never need another CTA's results: the CTA forms y_i from the k_j it wrote itself a
// stage earlier, evaluates the MLP, writes k_i and goes on -- no grid-wide synchronisation, one launch + ramp-up
// instead of six, TMEM and barriers set up once.  With 64-row tiles (mlp_tanh256_kernel<64, true>) a stage does not
// load its operands from global memory: the partial sum over all k_j but the newest is formed under the previous
// stage's MMAs and kept in shared memory, the newest k_j is read from the staged output tile, y from TMEM (see the
// comments in the kernel; DESIGN.md section 4 has the measurements).
constexpr int kMaxStageK = 6;
struct StageIn {
  float* k[kMaxStageK + 1];  // k[j], j < stage: operands; k[stage]: where stage `stage`'s f value goes
  float a[kMaxStageK + 1][kMaxStageK];  // a[i][j], j < i: row i of the tableau (data dtype)
  const void* dt;     // (B) per-sample step, float or double
  float* y_out;       // (B,256) or NULL: where y_i of the LAST stage is also stored (the step's y1)
  const int* ctl;     // control block or NULL: the launch is a no-op once the stop flag is set
  int stage0, stage1; // 0, 0: plain evaluation of y (no operands)
  int dt_is_f64;
};

// acc[0..8) = a[row][0] k_0, then fma(a[row][j], k_j, acc) for j < N, of the 8 elements at `off`: all 2 N loads first
template <int N>
__device__ __forceinline__ void partial_chunk(const StageIn& sp, int a_row, size_t off, float* acc) {
  float4 kv[N][2];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    kv[j][0] = *reinterpret_cast<const float4*>(sp.k[j] + off);
    kv[j][1] = *reinterpret_cast<const float4*>(sp.k[j] + off + 4);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float a = sp.a[a_row][j];
    const float kk[8] = {kv[j][0].x, kv[j][0].y, kv[j][0].z, kv[j][0].w, kv[j][1].x, kv[j][1].y, kv[j][1].z, kv[j][1].w};
#pragma unroll
    for (int x = 0; x < 8; ++x) acc[x] = j == 0 ? __fmul_rn(a, kk[x]) : __fmaf_rn(a, kk[x], acc[x]);
  }
}

// kStep (BM = 64 only): the launch covers stages 1 .. stage1 > 1 of a step -- operand rows through the partial sums
template <int kBM, bool kStep>
__global__ void __launch_bounds__(kThreads, 1)
mlp_tanh256_kernel(const float* __restrict__ y, const __nv_bfloat16* __restrict__ weights,
                   const float* __restrict__ biases, float* __restrict__ out, long long B, int n_layers,
                   const __grid_constant__ StageIn sp, const __grid_constant__ CUtensorMap wmap) {
  constexpr int kABlockBytes = kBM * 128;  // per K-block of the activation tile
  constexpr int kSmemA = smem_a_bytes(kBM);
  // 1024-byte alignment (SWIZZLE_128B atoms) is requested from the compiler / driver; no integer
  // round-up, so that the compiler keeps these pointers in the shared address space (LDS/STS)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kSmemA;
  float* sBias = reinterpret_cast<float*>(smem + kSmemA + kSmemW);    // BM = 128
  float4* sP = reinterpret_cast<float4*>(smem + kSmemA + kSmemW);     // BM = 64: [round][half][thread]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kSmemA + kSmemW + smem_extra_bytes(kBM));
  // mbar[0]: a layer's MMAs done, mbar[1]: its first half, mbar[2], mbar[3]: K-blocks 0-1 / 2-3 of a layer's weights
  // have landed (TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 4);
  constexpr uint32_t kTmemCols = tmem_cols(kStep);
  // step-fused: the staged fp32 output tile (64 KB) sits in K-blocks 2-3 of the weight buffer, where the next stage
  // reads its newest operand from; K-blocks 0-1 take the next stage's first weights meanwhile
  uint8_t* const sOutTile = sW + (kStep ? 2 * kWBlockBytes : 0);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * kBM;
#ifdef TODE_MLP_TIMING
  long long stamp[16];
  int n_stamp = 0;
  int stamp_stage = 0;  // -DTODE_MLP_TIMING=<first stage whose phases are stamped>
#define TODE_STAMP() do { if (tid == 0 && n_stamp < 16 && (n_stamp == 0 || stamp_stage >= TODE_MLP_TIMING)) stamp[n_stamp++] = clock64(); } while (0)
#ifdef TODE_MLP_TIMING_OPERAND  // extra stamps inside the operand phase, all taken by thread TODE_MLP_TIMING_OPERAND
#undef TODE_STAMP
#define TODE_STAMP() do { if (tid == TODE_MLP_TIMING_OPERAND && n_stamp < 16 && (n_stamp == 0 || stamp_stage >= TODE_MLP_TIMING)) stamp[n_stamp++] = clock64(); } while (0)
#define TODE_STAMP_X() TODE_STAMP()
#else
#define TODE_STAMP_X() do { } while (0)
#endif
#else
#define TODE_STAMP() do { } while (0)
#define TODE_STAMP_X() do { } while (0)
#endif
  TODE_STAMP();

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {  // the thread that issues the TMA loads and the MMAs
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar + 1)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar + 2)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar + 3)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // barriers initialised before the first TMA load is armed on one of them; the TMEM address is published
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // ---- layer 0 weights: asynchronous global -> shared copies, in flight during the A load ----
  // K-blocks 2 half, 2 half + 1 of a layer's weights (64 KB).  The weight buffer is refilled in halves: the first
  // two K-blocks as soon as the MMAs that read them are done -- under the second half of the layer's MMAs and the
  // epilogue -- the other two after the layer's last MMA (round 1 started the whole 128 KB only then and waited
  // ~0.8 k cycles for it at the top of the next layer, scripts/mlp_timing.py)
  const uint32_t wbar0 = smem_u32(mbar + 2), wbar1 = smem_u32(mbar + 3);
  uint32_t wparity = 0;
  // `after_generic`: the destination was last touched by ordinary loads / stores of this CTA (the staged
  // output tile): order them before the async-proxy writes of the TMA unit
  auto load_weights_half = [&](int layer, int half, bool after_generic = false) {
    if (tid == 0) {
      if (after_generic) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      const uint32_t wbar = half ? wbar1 : wbar0;
      mbar_arrive_expect_tx(wbar, 2 * kWBlockBytes);
#pragma unroll
      for (int kb = 2 * half; kb < 2 * half + 2; ++kb)
        tma_load_2d(smem_u32(sW) + kb * kWBlockBytes, &wmap, wbar, kb * kKBlock, layer * kWidth);
    }
  };
  auto load_weights_async = [&](int layer) {
    load_weights_half(layer, 0);
    load_weights_half(layer, 1);
  };
  load_weights_async(0);
  // every layer's bias, once (round 2: the per-layer load sat, with its full L2 latency, between a layer's epilogue
  // and the next layer's first MMA: ~0.6 k cycles per layer, scripts/mlp_timing.py)
  if constexpr (kBM == 128) {
    for (int idx = tid; idx < n_layers * kWidth; idx += kThreads) sBias[idx] = biases[idx];
  }

  // Everything above is independent of the kernel launched before this one (TMEM allocation,
  // barrier, the first layer's weights): under programmatic dependent launch it overlaps that
  // kernel's tail.  Its outputs -- y, k, the control block -- are only touched from here on.  The
  // next kernel may start its own prologue right away (it waits for OUR completion the same way).
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  if (sp.ctl != nullptr && sp.ctl[TODE_CTL_STOP]) {  // speculative iteration after the stop: undo the prologue
    if (tid == 0) {  // the first layer's weights are in flight: shared memory must outlive them
      mbar_wait(wbar0, 0);
      mbar_wait(wbar1, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp == 0) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    return;
  }

  // ---- step-fused: the step's y tile goes to TMEM once (32 columns of this thread's lane = its four 8-element
  // chunks), the rows' dt to registers; every stage of the step forms its rows from them
  const uint32_t y_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u + 32u * (uint32_t)(warp >> 2);
  float dtr[4] = {0.f, 0.f, 0.f, 0.f};
  if constexpr (kStep) {
    uint32_t yr[32];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = r * (kThreads / 32) + warp;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (m0 + row < B) {
        const float4* src = reinterpret_cast<const float4*>(y + (m0 + row) * kWidth + lane * 8);
        v0 = src[0];
        v1 = src[1];
        dtr[r] = sp.dt_is_f64 ? (float)static_cast<const double*>(sp.dt)[m0 + row]
                              : static_cast<const float*>(sp.dt)[m0 + row];
      }
      yr[r * 8 + 0] = __float_as_uint(v0.x), yr[r * 8 + 1] = __float_as_uint(v0.y);
      yr[r * 8 + 2] = __float_as_uint(v0.z), yr[r * 8 + 3] = __float_as_uint(v0.w);
      yr[r * 8 + 4] = __float_as_uint(v1.x), yr[r * 8 + 5] = __float_as_uint(v1.y);
      yr[r * 8 + 6] = __float_as_uint(v1.z), yr[r * 8 + 7] = __float_as_uint(v1.w);
    }
    tmem_st_32x32b_x32(y_taddr, yr);
  }

  uint32_t parity = 0;
  for (int stage = sp.stage0; stage <= sp.stage1; ++stage) {
#ifdef TODE_MLP_TIMING
  stamp_stage = stage;
  TODE_STAMP();  // stage begins
#endif
  const int nk = stage;                                    // operands k[0 .. nk-1]; 0: plain evaluation of y
  float* const outp = stage > 0 ? sp.k[stage] : out;       // where this evaluation's result goes
  float* const y_outp = stage == sp.stage1 ? sp.y_out : nullptr;
  // ---- activation tile: fp32 rows of y -> bf16, swizzled K-major (rows past B are zero) ----
  constexpr int kAChunks = kBM * (kWidth / 8);  // chunks of 8 elements
  constexpr int kAU = kAChunks / kThreads;      // chunks in flight per thread (2 x 16-byte loads each): 8 / 4
  // BM = 64, step-fused: no global loads on the way into a stage.  All operands but the newest were summed under the
  // previous stage's MMAs (sP, see partial_round below), the newest, k[nk-1], is still in the staged output tile of
  // the evaluation that produced it, y sits in TMEM, dt in registers.  (The first stage of the launch reads k[0],
  // which a previous launch produced, from global memory.)  Same chain: P = a_0 k_0, fma(a_j, k_j, P) ascending j,
  // the newest operand last, then fma(dt, acc, y).
  if (kStep && nk > 0) {
    constexpr int kRounds = kAChunks / kThreads;  // 4
    const float a_new = sp.a[stage][nk - 1];
    float4 kv[kRounds][2];
    if (nk == 1) {
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int row = r * (kThreads / 32) + warp;
        kv[r][0] = kv[r][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + row < B) {
          const float4* src = reinterpret_cast<const float4*>(sp.k[0] + (m0 + row) * kWidth + lane * 8);
          kv[r][0] = src[0];
          kv[r][1] = src[1];
        }
      }
    } else {
      const float* sOut = reinterpret_cast<const float*>(sOutTile);
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int row = r * (kThreads / 32) + warp;
        // (lanes 4-7 of every 8 take their second half first: the eight 16-byte loads of a quarter-warp then fall
        // into eight different bank groups instead of four)
        const float4* src = reinterpret_cast<const float4*>(sOut + row * kWidth + out_col<64>(row, lane * 8));
        const int swap = (lane >> 2) & 1;
        const float4 a = src[swap], b = src[swap ^ 1];
        kv[r][0] = swap ? b : a;
        kv[r][1] = swap ? a : b;
      }
    }
    TODE_STAMP_X();  // newest operand read
    uint32_t yr[32];
    tmem_ld_32x32b_x32(y_taddr, yr);
    TODE_STAMP_X();  // y read
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int row = r * (kThreads / 32) + warp, chunk = lane;
      float res[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) res[x] = 0.f;
      if (m0 + row < B) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nk > 1) pp = sP[(r * 2 + h) * kThreads + tid];
          const float kk[4] = {kv[r][h].x, kv[r][h].y, kv[r][h].z, kv[r][h].w};
          const float pa[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float acc = nk == 1 ? __fmul_rn(a_new, kk[x]) : __fmaf_rn(a_new, kk[x], pa[x]);
            res[h * 4 + x] = __fmaf_rn(dtr[r], acc, __uint_as_float(yr[r * 8 + h * 4 + x]));
          }
        }
        if (y_outp != nullptr) {
          const size_t off = (size_t)(m0 + row) * kWidth + chunk * 8;
          *reinterpret_cast<float4*>(y_outp + off) = make_float4(res[0], res[1], res[2], res[3]);
          *reinterpret_cast<float4*>(y_outp + off + 4) = make_float4(res[4], res[5], res[6], res[7]);
        }
      }
      uint4 p;
      p.x = pack_bf16(res[0], res[1]);
      p.y = pack_bf16(res[2], res[3]);
      p.z = pack_bf16(res[4], res[5]);
      p.w = pack_bf16(res[6], res[7]);
      *reinterpret_cast<uint4*>(sA + swz(kABlockBytes, chunk >> 3, row, chunk & 7)) = p;
    }
  } else if (!kStep && nk > 0) {
    // one 8-element chunk per thread and round; every operand row's two 16-byte loads are issued
    // before the first use (up to 14 loads in flight per thread)
#pragma unroll 1
    for (int base = 0; base < kAChunks; base += kThreads) {
      const int idx = base + tid;
      const int row = idx / (kWidth / 8), chunk = idx % (kWidth / 8);
      float r[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) r[x] = 0.f;
      if (m0 + row < B) {
        const size_t off = (size_t)(m0 + row) * kWidth + chunk * 8;
        float4 yv[2], kv[kMaxStageK][2];
        yv[0] = *reinterpret_cast<const float4*>(y + off);
        yv[1] = *reinterpret_cast<const float4*>(y + off + 4);
#pragma unroll
        for (int j = 0; j < kMaxStageK; ++j) {
          if (j < nk) {
            kv[j][0] = *reinterpret_cast<const float4*>(sp.k[j] + off);
            kv[j][1] = *reinterpret_cast<const float4*>(sp.k[j] + off + 4);
          }
        }
        const float dtr = sp.dt_is_f64 ? (float)static_cast<const double*>(sp.dt)[m0 + row]
                                       : static_cast<const float*>(sp.dt)[m0 + row];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float yy[4] = {yv[h].x, yv[h].y, yv[h].z, yv[h].w};
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxStageK; ++j) {
              if (j < nk) {
                const float4 kk = kv[j][h];
                const float kx = x == 0 ? kk.x : (x == 1 ? kk.y : (x == 2 ? kk.z : kk.w));
                acc = j == 0 ? __fmul_rn(sp.a[stage][0], kx) : __fmaf_rn(sp.a[stage][j], kx, acc);
              }
            }
            r[h * 4 + x] = __fmaf_rn(dtr, acc, yy[x]);
          }
        }
        if (y_outp != nullptr) {
          *reinterpret_cast<float4*>(y_outp + off) = make_float4(r[0], r[1], r[2], r[3]);
          *reinterpret_cast<float4*>(y_outp + off + 4) = make_float4(r[4], r[5], r[6], r[7]);
        }
      }
      uint4 p;
      p.x = pack_bf16(r[0], r[1]);
      p.y = pack_bf16(r[2], r[3]);
      p.z = pack_bf16(r[4], r[5]);
      p.w = pack_bf16(r[6], r[7]);
      *reinterpret_cast<uint4*>(sA + swz(kABlockBytes, chunk >> 3, row, chunk & 7)) = p;
    }
  } else {
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
  }

  TODE_STAMP_X();  // rows formed and stored
  // generic-proxy smem writes (st.shared) -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  TODE_STAMP_X();  // fenced
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  TODE_STAMP();  // activation tile loaded
  // step-fused, later stages: the staged output tile has been read; K-blocks 2-3 of this stage's first layer take
  // its place (they are needed half-way through the layer's MMAs)
  if (kStep && stage > sp.stage0) load_weights_half(0, 1, true);
  const uint32_t idesc = make_idesc(kBM, kWidth);
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW), bar = smem_u32(mbar), bar_half = smem_u32(mbar + 1);

  // Round `r` (512 chunks of 8 elements) of the NEXT stage's partial operand sum, run by the threads while they
  // would otherwise wait for this stage's MMAs: P = a[s+1][0] k_0, fma(a[s+1][j], k_j, P) for j < s (all of them
  // written before this stage began; k_s itself joins when the next stage forms its rows).  Every thread reads
  // back only what it wrote: no barrier.
  const bool do_partial = kStep && stage >= 1 && stage < sp.stage1;
  auto partial_round = [&](int r) {
    const int idx = r * kThreads + tid;
    const int row = idx / (kWidth / 8), chunk = idx % (kWidth / 8);
    float acc[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) acc[x] = 0.f;
    if (m0 + row < B) {
      const size_t off = (size_t)(m0 + row) * kWidth + chunk * 8;
      switch (stage) {  // one fully unrolled body per operand count: everything stays in registers
        case 1: partial_chunk<1>(sp, stage + 1, off, acc); break;
        case 2: partial_chunk<2>(sp, stage + 1, off, acc); break;
        case 3: partial_chunk<3>(sp, stage + 1, off, acc); break;
        case 4: partial_chunk<4>(sp, stage + 1, off, acc); break;
        default: partial_chunk<5>(sp, stage + 1, off, acc); break;
      }
    }
    sP[(r * 2 + 0) * kThreads + tid] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    sP[(r * 2 + 1) * kThreads + tid] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  };

  for (int layer = 0; layer < n_layers; ++layer) {
    // ---- this layer's weights (out, in) = (N, K) row-major -> K-major swizzled, and bias ------
    // (issued right after the previous layer's MMAs completed, in flight during its epilogue)
    // (the activation tile is complete and fenced: barrier after the operand rows / the previous epilogue)
    if (warp == 0 && lane == 0) {
      mbar_wait(wbar0, wparity);  // K-blocks 0-1 of this layer's weights have landed
      TODE_STAMP();               // (first half of) this layer's weights have arrived
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int kb = 0; kb < kNumKBlocks; ++kb) {
        if (kb == kNumKBlocks / 2) mbar_wait(wbar1, wparity);  // K-blocks 2-3
#pragma unroll
        for (int ks = 0; ks < kKBlock / 16; ++ks) {
          const uint64_t a_desc = make_desc(sA_addr + kb * kABlockBytes + ks * 32);
          const uint64_t b_desc = make_desc(sW_addr + kb * kWBlockBytes + ks * 32);
          mma_bf16(tmem_base, a_desc, b_desc, idesc, (kb | ks) != 0 ? 1u : 0u);
        }
        if (kb == kNumKBlocks / 2 - 1) {  // the tensor core is done with K-blocks 0, 1 of the weight tile
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_half)
                       : "memory");
        }
      }
      // arrives on the mbarrier when every MMA above has completed (implies fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
                   : "memory");
    }
    wparity ^= 1;
    const bool last_layer = layer == n_layers - 1;
    // step-fused: the last layer's output tile is staged in K-blocks 2-3, so K-blocks 0-1 of the NEXT STAGE's first
    // layer can travel as soon as this layer's MMAs are done with theirs
    const bool next_stage_early = kStep && last_layer && stage < sp.stage1;
    // rounds of the next stage's partial sum under this layer's MMAs (L2 -> SM bandwidth bounds them: ~2.4 k cycles
    // each with the weights streaming beside them): one each under layers 0 and 1, two under layer 2, where no next
    // layer's weights travel (1 layer: all four; 2 layers: two each); the first of them before the half-way barrier
    int pr = 4, pr_end = 4;
    if (kStep && do_partial && layer < 3) {
      if (n_layers == 1) {
        pr = 0, pr_end = 4;
      } else if (n_layers == 2) {
        pr = 2 * layer, pr_end = layer < 2 ? pr + 2 : pr;
      } else {
        pr = layer, pr_end = layer == 2 ? 4 : layer + 1;
      }
      if (pr < pr_end) partial_round(pr++);
    }
    mbar_wait(bar_half, parity);
    if (!last_layer) load_weights_half(layer + 1, 0);
    if (next_stage_early) load_weights_half(0, 0);
    if constexpr (kStep) {
#pragma unroll 1
      for (; pr < pr_end; ++pr) partial_round(pr);
    }
    // BM = 64: this thread's 16 bias values of the layer travel under the MMAs
    // (the epilogue's thread coordinates come from a laundered copy of tid: its ~40 loop-invariant addresses are
    // then formed here, per layer, instead of living in registers across the operand rows and the partial rounds)
    int tid_e = tid;
    asm volatile("" : "+r"(tid_e));
    const int warp_e = tid_e >> 5, lane_e = tid_e & 31;
    float2 bia[8];
    if constexpr (kBM == 64) {
      const float2* bg = reinterpret_cast<const float2*>(biases + layer * kWidth + (warp_e >> 2) * 64 + 2 * (lane_e & 3));
#pragma unroll
      for (int j = 0; j < 8; ++j) bia[j] = bg[4 * j];
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    TODE_STAMP();  // MMAs done
    // the tensor core is done reading sW: fetch the next layer's weights behind the epilogue
    if (!last_layer) load_weights_half(layer + 1, 1);

    // ---- epilogue: TMEM -> registers, + bias, tanh, -> next layer's activation tile / out ----
    const bool last = layer == n_layers - 1;
    const int q = warp_e & 3, cq = warp_e >> 2;  // TMEM lane quarter, column quarter
    if constexpr (kBM == 128) {
      epilogue_m128(tmem_base, sA, sW, sBias + layer * kWidth, q, cq, lane_e, last);
    } else {
      epilogue_m64(tmem_base, sA, sOutTile, bia, q, cq, lane_e, last);
    }
    if (last) {
      __syncthreads();
      // coalesced copy-out: one warp writes whole 1 KB rows, 128 contiguous bytes per instruction
      const float* sOut = reinterpret_cast<const float*>(sOutTile);
      for (int r = warp; r < kBM; r += kThreads / 32) {
        if (m0 + r < B) {
          float* dst = outp + (m0 + r) * kWidth;
          if constexpr (kBM == 64) {  // the row's rotation keeps 4-column groups: 16-byte reads and stores
#pragma unroll
            for (int j = 0; j < kWidth / 128; ++j) {
              const int col = 4 * (lane + 32 * j);
              *reinterpret_cast<float4*>(dst + col) =
                  *reinterpret_cast<const float4*>(sOut + r * kWidth + out_col<64>(r, col));
            }
          } else {
#pragma unroll
            for (int j = 0; j < kWidth / 32; ++j) dst[lane + 32 * j] = sOut[r * kWidth + out_col<kBM>(r, lane + 32 * j)];
          }
        }
      }
    }
    // TMEM reads and smem writes of this layer are done before the next layer's MMA starts
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    TODE_STAMP();  // epilogue done
  }
  // next stage of a multi-stage launch with 128-row tiles (the output tile took the whole weight buffer): its first
  // layer's weights travel while its operand rows are loaded
  if (!kStep && stage < sp.stage1) {
    load_weights_half(0, 0, true);
    load_weights_half(0, 1, true);
  }
  }  // stage
#ifdef TODE_MLP_TIMING
#ifdef TODE_MLP_TIMING_OPERAND
  const int stamp_tid = TODE_MLP_TIMING_OPERAND;
#else
  const int stamp_tid = 0;
#endif
  __syncthreads();
  if (tid == stamp_tid && blockIdx.x == 0) {
    for (int i = 0; i < n_stamp; ++i) reinterpret_cast<long long*>(out)[i] = stamp[i] - stamp[0];
    reinterpret_cast<long long*>(out)[n_stamp] = -1;
  }
#endif

  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace mlp
}  // namespace tode

namespace tode {
namespace mlp {

// programmatic dependent launch of the stage-fused evaluations; TORCHODE_B200_PDL=0 turns it off
static bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("TORCHODE_B200_PDL");
    return v == nullptr ? TODE_MLP_PDL_DEFAULT != 0 : (v[0] != '0');
  }();
  return on;
}
static int launch_error() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// Tiled tensor map over the weights viewed as (n_layers * 256 rows, 256 columns) bf16: boxes of 256 rows x 64
// columns (one K-block of one layer, 32 KB) land in shared memory in the SWIZZLE_128B layout.  Encoding is a
// pure host computation (~1 us): done per launch, nothing cached, nothing to go stale.
static int weight_tensor_map(const void* weights_bf16, int n_layers, CUtensorMap* map) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (encode == nullptr) return TODE_ENOSUP;
  const cuuint64_t dims[2] = {(cuuint64_t)kWidth, (cuuint64_t)kWidth * (cuuint64_t)n_layers};
  const cuuint64_t strides[1] = {(cuuint64_t)kWidth * 2};  // bytes between rows
  const cuuint32_t box[2] = {(cuuint32_t)kKBlock, (cuuint32_t)kWidth};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(weights_bf16), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : TODE_EINVAL;
}

static int launch_mlp(const float* y, const void* weights_bf16, const void* biases_f32, void* out, int64_t B,
                      int32_t n_layers, const StageIn& sp, void* stream) {
  if (!y || !weights_bf16 || !biases_f32 || !out || n_layers < 1 || n_layers > kMaxLayers) return TODE_EINVAL;
  if (B == 0) return 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(y) || !al16(weights_bf16) || !al16(out) || !al16(biases_f32)) return TODE_EALIGN;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_tanh256_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes(128));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp_tanh256_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(64));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp_tanh256_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(64));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // 128-row tiles unless they would leave SMs idle while 64-row tiles fill more of them
  const bool big = (B + 127) / 128 >= sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(big ? (B + 127) / 128 : (B + 63) / 64));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes(big ? 128 : 64);
  cfg.stream = static_cast<cudaStream_t>(stream);
  // stage-fused evaluations follow a kernel of the same solve in the stream: let this launch's
  // prologue overlap that kernel's tail (the kernel waits with griddepcontrol.wait before it
  // reads anything the predecessor wrote)
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (sp.stage0 > 0 && pdl_enabled()) ? 1 : 0;
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(weights_bf16);
  const float* bias = static_cast<const float*>(biases_f32);
  float* o = static_cast<float*>(out);
  const long long rows = (long long)B;
  const int layers = (int)n_layers;
  CUtensorMap wmap{};
  if (const int rc = weight_tensor_map(weights_bf16, layers, &wmap)) return rc;
  const bool step = !big && sp.stage0 == 1 && sp.stage1 > 1;
  const cudaError_t e = big ? cudaLaunchKernelEx(&cfg, mlp_tanh256_kernel<128, false>, y, w, bias, o, rows, layers, sp, wmap)
                        : step ? cudaLaunchKernelEx(&cfg, mlp_tanh256_kernel<64, true>, y, w, bias, o, rows, layers, sp, wmap)
                               : cudaLaunchKernelEx(&cfg, mlp_tanh256_kernel<64, false>, y, w, bias, o, rows, layers, sp, wmap);
  if (e != cudaSuccess) return (int)e;
  return launch_error();
}

}  // namespace mlp
}  // namespace tode

extern "C" int tode_mlp_tanh256_forward(const void* y, const void* weights_bf16, const void* biases_f32,
                                        void* out, int64_t B, int32_t n_layers, void* stream) {
  tode::mlp::StageIn sp{};
  return tode::mlp::launch_mlp(static_cast<const float*>(y), weights_bf16, biases_f32, out, B, n_layers, sp, stream);
}

extern "C" int tode_mlp_tanh256_stage_forward(const tode_tableau* tab, int stage, const tode_state* st,
                                              const void* const* k, void* y_out, const void* weights_bf16,
                                              const void* biases_f32, void* out, int32_t n_layers, void* stream) {
  using namespace tode::mlp;
  if (!tab || !st || !k || !st->y || !st->dt) return TODE_EINVAL;
  if (stage < 1 || stage >= tab->n_stages || stage > kMaxStageK) return TODE_EINVAL;
  if (st->F != kWidth || st->data_dtype != TODE_F32) return TODE_ENOSUP;
  if (st->time_dtype != TODE_F32 && st->time_dtype != TODE_F64) return TODE_EINVAL;
  StageIn sp{};
  for (int j = 0; j < stage; ++j) {
    if (!k[j]) return TODE_EINVAL;
    if (reinterpret_cast<uintptr_t>(k[j]) & 15) return TODE_EALIGN;
    sp.k[j] = const_cast<float*>(static_cast<const float*>(k[j]));
    sp.a[stage][j] = (float)tab->a[stage][j];  // ButcherTableau.to(data dtype), as tode_erk_stage
  }
  if (!out || (reinterpret_cast<uintptr_t>(out) & 15)) return out ? TODE_EALIGN : TODE_EINVAL;
  sp.k[stage] = static_cast<float*>(out);
  if (y_out && (reinterpret_cast<uintptr_t>(y_out) & 15)) return TODE_EALIGN;
  sp.dt = st->dt;
  sp.y_out = static_cast<float*>(y_out);
  sp.ctl = st->ctl;
  sp.stage0 = sp.stage1 = stage;
  sp.dt_is_f64 = st->time_dtype == TODE_F64;
  return launch_mlp(static_cast<const float*>(st->y), weights_bf16, biases_f32, out, st->B, n_layers, sp, stream);
}

extern "C" int tode_mlp_tanh256_step_forward(const tode_tableau* tab, const tode_state* st, void* const* k,
                                             void* y1_out, const void* weights_bf16, const void* biases_f32,
                                             int32_t n_layers, void* stream) {
  using namespace tode::mlp;
  if (!tab || !st || !k || !st->y || !st->dt) return TODE_EINVAL;
  if (tab->n_stages < 2 || tab->n_stages - 1 > kMaxStageK) return TODE_ENOSUP;
  if (st->F != kWidth || st->data_dtype != TODE_F32) return TODE_ENOSUP;
  if (st->time_dtype != TODE_F32 && st->time_dtype != TODE_F64) return TODE_EINVAL;
  StageIn sp{};
  for (int i = 0; i < tab->n_stages; ++i) {
    if (!k[i]) return TODE_EINVAL;
    if (reinterpret_cast<uintptr_t>(k[i]) & 15) return TODE_EALIGN;
    sp.k[i] = static_cast<float*>(k[i]);
    for (int j = 0; j < i; ++j) sp.a[i][j] = (float)tab->a[i][j];
  }
  if (y1_out && (reinterpret_cast<uintptr_t>(y1_out) & 15)) return TODE_EALIGN;
  sp.dt = st->dt;
  sp.y_out = static_cast<float*>(y1_out);
  sp.ctl = st->ctl;
  sp.stage0 = 1;
  sp.stage1 = tab->n_stages - 1;
  sp.dt_is_f64 = st->time_dtype == TODE_F64;
  return launch_mlp(static_cast<const float*>(st->y), weights_bf16, biases_f32, sp.k[sp.stage1], st->B, n_layers, sp,
                    stream);
}
