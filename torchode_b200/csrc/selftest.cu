// C-ABI: tode_selftest_fast_math -- test aid (tests/test_gpu_kernels.py), not on the solve path.
// Compares the branch-free scalar functions of the fused kernel (erk_math.cuh "fast path")
// with the checked ones on pseudo-random and adversarial operands: wherever the fast
// function leaves its range flag set, the two must agree bit for bit.
#include "api_common.cuh"
#include "erk_fused_f2.cuh"  // the fp32 fast-path division / square root of the packed kernels

namespace tode {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// operand generator: mode 0 = any bit pattern; 1 = moderate exponents (the solver's usual
// range); 2 = near the exponent-range limits / special values
__device__ double gen(unsigned long long h, int mode) {
  if (mode == 0) return __longlong_as_double((long long)h);
  const unsigned long long mant = h & 0x000fffffffffffffull;
  const unsigned long long sign = (h >> 63) << 63;
  if (mode == 1) {
    const long long e = 1023 - 40 + (long long)((h >> 52) % 80);
    return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
  }
  const unsigned sel = (unsigned)((h >> 52) & 15);
  switch (sel) {
    case 0: return 0.0;
    case 1: return -0.0;
    case 2: return __longlong_as_double(0x7ff0000000000000LL);
    case 3: return __longlong_as_double((long long)0xfff0000000000000ull);
    case 4: return __longlong_as_double(0x7ff8000000000000LL);
    case 5: return 1.0;
    case 6: return __longlong_as_double((long long)(sign | mant));                          // subnormal
    case 7: return __longlong_as_double((long long)(sign | (0x7feull << 52) | mant));       // huge
    case 8: return __longlong_as_double((long long)(sign | (1ull << 52) | mant));           // tiny normal
    case 9: return __longlong_as_double((long long)(sign | ((1023ull - 969) << 52) | mant));
    case 10: return __longlong_as_double((long long)(sign | ((1023ull + 1017) << 52) | mant));
    case 11: return __longlong_as_double((long long)(sign | ((1023ull - 1) << 52) | 0xfffffffffffffull));
    case 12: return __longlong_as_double((long long)(sign | (1023ull << 52) | (mant & 7)));  // 1 + few ulp
    default: {
      const long long e = 1 + (long long)((h >> 40) % 2046);
      return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
    }
  }
}

__device__ __forceinline__ bool same_bits(double a, double b) {
  return __double_as_longlong(a) == __double_as_longlong(b);
}

// counts[0..3]: mismatches of division / log2 / exp2 / controller; counts[4..7]: how often the
// fast flag stayed set (so a test can tell that the fast path was actually exercised)
__global__ void selftest_kernel(long long n, unsigned long long seed, CtrlP<double, double> c,
                                const __grid_constant__ PowTab pt, unsigned long long* counts) {
  __shared__ __align__(16) double s_pow[kPowSharedDoubles];
  pow_tables_to_shared(s_pow, threadIdx.x, blockDim.x);
  __syncthreads();
  const PowShared ps{pt, reinterpret_cast<const double2*>(s_pow)};
  unsigned long long bad[4] = {0, 0, 0, 0}, used[4] = {0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h0 = mix64(seed + 4 * (unsigned long long)i);
    const unsigned long long h1 = mix64(h0), h2 = mix64(h1), h3 = mix64(h2);
    const int mode = (int)(i % 3);
    const double a = gen(h0, mode), b = gen(h1, mode);
    {  // division (single, paired, by a fixed divisor)
      bool ok = true;
      const double q = div_chk(a, b, ok);
      if (ok) { used[0]++; if (!same_bits(q, __ddiv_rn(a, b))) bad[0]++; }
      const double av[2] = {a, gen(h2, mode)}, bv[2] = {b, gen(h3, 1)};
      double qv[2];
      bool ok2 = true;
      div_chk_n<2>(av, bv, qv, ok2);
      if (ok2 && (!same_bits(qv[0], __ddiv_rn(av[0], bv[0])) || !same_bits(qv[1], __ddiv_rn(av[1], bv[1])))) bad[0]++;
      const DivBy<double> by(sqrt((double)(1 + (int)(h3 & 7))));
      bool ok3 = true;
      const double q3 = by(a, ok3);
      if (ok3 && !same_bits(q3, __ddiv_rn(a, by.c))) bad[0]++;
    }
    {  // log2
      bool ok = true;
      const double x = fabs(a);
      const double l = det_log2_fast(x, ok, ps);
      if (ok) { used[1]++; if (!same_bits(l, det_log2_safe(x))) bad[1]++; }
    }
    {  // exp2 (arguments mostly in the useful range)
      const double z[3] = {mode == 1 ? a * 1e-9 : a, ldexp(gen(h2, 1), -30), gen(h3, mode)};
      double p[3];
      bool ok = true;
      det_exp2_fast<3>(z, p, ok, ps);
      if (ok) {
        used[2]++;
        for (int j = 0; j < 3; ++j)
          if (!same_bits(p[j], det_exp2(z[j]))) bad[2]++;
      }
    }
    {  // controller: norm anywhere, history = 1 or accepted ratios in [almost_zero, 1)
      const double nrm = (mode == 1) ? fabs(a) * ldexp(1.0, (int)(h2 % 40) - 20) : fabs(a);
      double r1 = 1.0, r2 = 1.0;
      if (h3 & 1) r1 = fmax(c.almost_zero, ldexp(1.0 + (double)(h3 >> 12) * 0x1p-52, -1 - (int)((h3 >> 2) % 60)));
      if (h3 & 2) r2 = fmax(c.almost_zero, ldexp(1.0 + (double)(h2 >> 12) * 0x1p-52, -1 - (int)((h2 >> 2) % 60)));
      const double L1 = det_log2_safe(r1), L2 = det_log2_safe(r2);
      const double dt = gen(h1, 1);
      CtrlP<double, double> cc = c;
      cc.pid = (int)((h2 >> 20) & 1);
      if ((h2 >> 21) & 1) cc.e_prev2 = 0.0;
      bool ok = true;
      const CtrlOut<double, double> f = controller_fast<double, double>(cc, nrm, dt, r1, r2, L1, L2, ok, ps);
      if (ok) {
        used[3]++;
        double Lr;
        const CtrlOut<double, double> s = controller_l<double, double>(cc, nrm, dt, r1, r2, L1, L2, &Lr);
        if (!same_bits(f.dt_next, s.dt_next) || !same_bits(f.ratio, s.ratio) || !same_bits(f.r1, s.r1) ||
            !same_bits(f.r2, s.r2) || !same_bits(f.L_ratio, s.L_ratio) || f.status != s.status ||
            f.accept != s.accept)
          bad[3]++;
      }
    }
  }
  for (int j = 0; j < 4; ++j) {
    if (bad[j]) atomicAdd(&counts[j], bad[j]);
    if (used[j]) atomicAdd(&counts[4 + j], used[j]);
  }
}

}  // namespace
}  // namespace tode

extern "C" int tode_selftest_fast_math(int64_t n, uint64_t seed, const tode_controller* ctrl, void* counts8,
                                       void* stream) {
  if (n <= 0 || !ctrl || !counts8) return TODE_EINVAL;
  const tode::CtrlP<double, double> c = tode::make_ctrl<double, double>(ctrl);
  tode::selftest_kernel<<<tode::sm_count() * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (long long)n, (unsigned long long)seed, c, tode::make_powtab(), static_cast<unsigned long long*>(counts8));
  return tode::launch_status();
}

// ---- fp32: div_fast / div_fast2 / sqrt_fast of erk_fused_f2.cuh against div.rn.f32 / sqrt.rn.f32 ------------
namespace tode {
namespace {

__device__ float gen32(unsigned long long h, int mode) {
  const unsigned int w = (unsigned int)(h >> 17);
  if (mode == 0) return __uint_as_float(w);  // any bit pattern
  const unsigned int mant = w & 0x007fffffu, sign = w & 0x80000000u;
  if (mode == 1) return __uint_as_float(sign | ((127u - 30u + (unsigned int)((h >> 8) % 60)) << 23) | mant);  // the solver's range
  // around the limits of mid_range ([2^-60, 2^60]) and of sqrt's fast path (2^-101), zeros, subnormals
  switch ((unsigned int)(h & 7)) {
    case 0: return __uint_as_float(sign);
    case 1: return __uint_as_float(sign | mant);
    case 2: return __uint_as_float(sign | ((127u - 61u + (unsigned int)((h >> 8) % 3)) << 23) | mant);
    case 3: return __uint_as_float(sign | ((127u + 59u + (unsigned int)((h >> 8) % 3)) << 23) | mant);
    case 4: return __uint_as_float(((127u - 102u + (unsigned int)((h >> 8) % 3)) << 23) | mant);
    case 5: return __uint_as_float(sign | (0x7f000000u) | mant);
    case 6: return __uint_as_float(sign | 0x3f800000u | (mant & 3));
    default: return __uint_as_float(sign | ((1u + (unsigned int)((h >> 8) % 253)) << 23) | mant);
  }
}

// counts[0..2]: mismatches of division (scalar and packed, shared reciprocal) / division by sqrt(2) / square
// root; counts[3..5]: how often the fast path's range flag was set
__global__ void selftest_f32_kernel(long long n, unsigned long long seed, unsigned long long* counts) {
  unsigned long long bad[3] = {0, 0, 0}, used[3] = {0, 0, 0};
  const float sqrt2 = (float)sqrt(2.0), r_sqrt2 = rcp_refined(sqrt2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h0 = mix64(seed + 3 * (unsigned long long)i), h1 = mix64(h0), h2 = mix64(h1);
    const int mode = (int)(i % 3);
    const float a = gen32(h0, mode), b = gen32(h1, mode), a2 = gen32(h2, mode);
    if (mid_range(b)) {
      const float r = rcp_refined(b);
      if (mid_range(a) || a == 0.0f) {  // (the kernels pass |err| or a positive time difference: +0, never -0)
        used[0]++;
        const float fa = a == 0.0f ? 0.0f : a;
        if (__float_as_uint(div_fast(fa, b, r)) != __float_as_uint(__fdiv_rn(fa, b))) bad[0]++;
      }
      if (mid_range(a) && mid_range(a2)) {  // two numerators, one divisor (the t_eval loop), packed
        const float2 q = div_fast2(make_float2(a, a2), splat(b), splat(r));
        if (__float_as_uint(q.x) != __float_as_uint(__fdiv_rn(a, b)) ||
            __float_as_uint(q.y) != __float_as_uint(__fdiv_rn(a2, b)))
          bad[0]++;
      }
      if (mid_range(a) && mid_range(a2)) {  // two divisions with their own divisors (|err| / bounds), packed
        const float b2 = fabsf(a2);
        const float2 q = div_fast2(make_float2(fabsf(a), fabsf(b)), make_float2(fabsf(b), b2),
                                   rcp_refined2(make_float2(fabsf(b), b2)));
        if (__float_as_uint(q.x) != __float_as_uint(__fdiv_rn(fabsf(a), fabsf(b))) ||
            __float_as_uint(q.y) != __float_as_uint(__fdiv_rn(fabsf(b), b2)))
          bad[0]++;
      }
    }
    if (mid_range(a)) {
      used[1]++;
      if (__float_as_uint(div_fast(a, sqrt2, r_sqrt2)) != __float_as_uint(__fdiv_rn(a, sqrt2))) bad[1]++;
    }
    {
      const float x = fabsf(a);
      bool ok = true;
      const float s = sqrt_fast(x, ok);
      if (ok) {
        used[2]++;
        if (__float_as_uint(s) != __float_as_uint(__fsqrt_rn(x))) bad[2]++;
      }
    }
  }
  for (int j = 0; j < 3; ++j) {
    if (bad[j]) atomicAdd(&counts[j], bad[j]);
    if (used[j]) atomicAdd(&counts[3 + j], used[j]);
  }
}

}  // namespace
}  // namespace tode

extern "C" int tode_selftest_fast_math_f32(int64_t n, uint64_t seed, void* counts6, void* stream) {
  if (n <= 0 || !counts6) return TODE_EINVAL;
  tode::selftest_f32_kernel<<<tode::sm_count() * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (long long)n, (unsigned long long)seed, static_cast<unsigned long long*>(counts6));
  return tode::launch_status();
}
