// Probe: per-SM throughput of MUFU.TANH against MUFU.EX2 / MUFU.RCP / FFMA (sm_100a), 16 warps per SM as in the
// MLP field's epilogue.  Prints lanes per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_rates mufu_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) rate_kernel(float* out, long long* cycles, int iters) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + 37 * i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (OP == 4) {  // half tanh, half 4 FFMA (do the pipes overlap?)
        if (i & 1) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
        else asm volatile("fma.rn.f32 %0, %0, %0, %0;\n fma.rn.f32 %0, %0, %0, %0;\n fma.rn.f32 %0, %0, %0, %0;\n fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double ops_per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  rate_kernel<OP><<<148, 512>>>(out, cyc, iters);
  rate_kernel<OP><<<148, 512>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s %8.2f lanes/clk/SM (%lld cycles)\n", name, ops_per_iter * 512.0 * iters / (double)h[0], h[0]);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("MUFU.TANH", 8);
  run<1>("MUFU.EX2", 8);
  run<2>("MUFU.RCP", 8);
  run<3>("FFMA", 8);
  run<4>("4 TANH + 16 FFMA (ops = 20)", 20);
  return 0;
}
