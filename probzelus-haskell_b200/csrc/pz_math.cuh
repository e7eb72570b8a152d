// pz_math.cuh -- device side of the canonical math spec (DESIGN.md "Canonical math" and "RNG").
//
// Counter-based Philox4x32-10 replaces the reference's sequential MWC256 SamplerIO
// (monad-bayes Control.Monad.Bayes.Sampler; call sites Inference.hs:30-42): a draw is a pure
// function of (seed, global slot, step, call), so particles are reproducible independently
// of launch shape and GPU count.
//
// Every fp32 operation here is an individually rounded IEEE op (the translation unit is
// compiled with -fmad=false); fused multiply-adds appear exactly where __fmaf_rn is
// written.  The CPU oracle evaluates the same expressions on the same constants
// (include/pz_spec_tables.h), which is what makes whole-filter trajectories bit-comparable.
// No MUFU approximation, conversion or division appears on the hot path: a standard normal is
// one table row + 3 fmaf, a weight is 5 fmaf + integer exponent arithmetic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/pz_spec_tables.h"

namespace pz {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

constexpr uint32_t kStreamProp = 0x50524F50u;      // propagation noise: ctr = (group, call, step, "PROP")
constexpr uint32_t kStreamResample = 0x52455341u;  // the one uniform of a systematic resample
constexpr uint32_t kStreamMulti = 0x4D554C54u;     // per-slot uniforms of the multinomial resampler

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
#else
    uint64_t p0 = (uint64_t)kPhiloxM0 * c0, p1 = (uint64_t)kPhiloxM1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += kPhiloxW0; k1 += kPhiloxW1;
  }
  return U4{c0, c1, c2, c3};
}

constexpr float kLog2e = 0x1.715476p+0f;
constexpr int32_t kNZero = -(1 << 30);  // exponent marker of an empty block (every weight zero)
constexpr int kGuard = 16;              // guard bits of the global fixed-point scale

// Inverse normal CDF table (256 rows of 4 coefficients); the step kernel stages it in shared memory.
static __device__ const float4 g_icdf_tab[PZ_ICDF_ROWS] = PZ_ICDF_TABLE_INIT;

// One 32-bit word -> one N(0,1) draw.  Bit 31 is the sign; v = (r << 1) | 1 is the tail
// probability u = v * 2^-33; clz(v) selects the dyadic segment, the next 3 bits the sub-segment,
// the following 23 bits the cubic's argument tf in [1, 2).  tab_m8 = table base - 8 rows (the
// sub-segment field carries the leading one).
__device__ __forceinline__ float icdf_normal(uint32_t r, const float4* __restrict__ tab_m8) {
  uint32_t v, k;
  asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(v) : "r"(r));          // (r << 1) | 1 in one IMAD
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(k) : "r"(v));        // clz(v) (v != 0): FLO.SH, no 31 - x
  const uint32_t vn = v << k;
  const float4 c = tab_m8[(k << PZ_ICDF_SUB_BITS) + (vn >> 28)];
  const float tf = __uint_as_float(0x3F800000u | ((vn >> 5) & 0x7FFFFFu));
  const float z = __fmaf_rn(__fmaf_rn(__fmaf_rn(c.w, tf, c.z), tf, c.y), tf, c.x);
  return __uint_as_float(__float_as_uint(z) ^ (r & 0x80000000u));
}

// Block exponent n_b (as float and int) from the block's largest log-weight; kNZero when empty.
__device__ __forceinline__ float block_exponent(float maxlw, int32_t& nb) {
  if (!(maxlw > -INFINITY)) { nb = kNZero; return 0.0f; }
  const float e = fminf(fmaxf(__fmul_rn(maxlw, kLog2e), -262144.0f), 262144.0f);
  const float nbf = rintf(e);
  nb = __float2int_rn(nbf);
  return nbf;
}

// Weight relative to 2^n_b and its fixed-point image: w = 2^d, d = max(lw * log2(e) - n_b, -32);
// qbits = 0x4B000000 + rint(w * 2^15) (q <= 46990 sits in the low 16 bits).  CLAMP_HI adds the
// upper clamp of the spec, unreachable when lw <= the block maximum and |n_b| < 2^18 (fused path).
// weight_q_from_arg takes d0 = lw * log2(e) - n_b already rounded to fp32 (an f64 filter forms it in double).
template <bool CLAMP_HI>
__device__ __forceinline__ void weight_q_from_arg(float d0, float& w, uint32_t& qbits) {
  float d = fmaxf(d0, -32.0f);
  if (CLAMP_HI) d = fminf(d, 1.0f);
  const float t = __fadd_rn(d, 12582912.0f);
  const float r = __fsub_rn(t, 12582912.0f);
  const float f = __fsub_rn(d, r);
  float p = PZ_EXP2_C5;
  p = __fmaf_rn(p, f, PZ_EXP2_C4);
  p = __fmaf_rn(p, f, PZ_EXP2_C3);
  p = __fmaf_rn(p, f, PZ_EXP2_C2);
  p = __fmaf_rn(p, f, PZ_EXP2_C1);
  p = __fmaf_rn(p, f, 1.0f);
  w = __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
  qbits = __float_as_uint(__fmaf_rn(w, 32768.0f, 8388608.0f));
}
template <bool CLAMP_HI>
__device__ __forceinline__ void weight_q(float lw, float nbf, float& w, uint32_t& qbits) {
  weight_q_from_arg<CLAMP_HI>(__fmaf_rn(lw, kLog2e, -nbf), w, qbits);
}

// Two weights at once on the packed fp32 pipe (FFMA2 / FADD2 / FMUL2 of sm_100: two independent, individually
// rounded IEEE operations per instruction, so the bits are those of weight_q<false> applied to each half; the step
// kernel is instruction-issue bound and this halves the issue slots of the polynomial).
__device__ __forceinline__ float2 pk(float a) { return make_float2(a, a); }
__device__ __forceinline__ void weight_q2(float2 lw, float nbf, float2& w, uint32_t& qb0, uint32_t& qb1) {
  float2 d = __ffma2_rn(lw, pk(kLog2e), pk(-nbf));
  d.x = fmaxf(d.x, -32.0f); d.y = fmaxf(d.y, -32.0f);
  const float2 t = __fadd2_rn(d, pk(12582912.0f));
  const float2 r = __fadd2_rn(t, pk(-12582912.0f));
  const float2 f = __fadd2_rn(d, make_float2(-r.x, -r.y));
  float2 p = pk(PZ_EXP2_C5);
  p = __ffma2_rn(p, f, pk(PZ_EXP2_C4));
  p = __ffma2_rn(p, f, pk(PZ_EXP2_C3));
  p = __ffma2_rn(p, f, pk(PZ_EXP2_C2));
  p = __ffma2_rn(p, f, pk(PZ_EXP2_C1));
  p = __ffma2_rn(p, f, pk(1.0f));
  w.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
  w.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
  const float2 q = __ffma2_rn(w, pk(32768.0f), pk(8388608.0f));
  qb0 = __float_as_uint(q.x); qb1 = __float_as_uint(q.y);
}

// stored log-weights may hold anything: +inf and NaN count as -inf (zero weight)
__device__ __forceinline__ float sanitize_lw(float lw) { return (lw == lw && lw < INFINITY) ? lw : -INFINITY; }

// block sum at the global scale: s * 2^(kGuard - ksh), truncated
__device__ __forceinline__ uint64_t shift_guard(uint64_t s, int32_t ksh) {
  if (ksh <= kGuard) return s << (kGuard - ksh);
  const int32_t sh = ksh - kGuard;
  return sh >= 64 ? 0ull : (s >> sh);
}

// memory offset inside a canonical block of the particle at CDF position p (see resample_tile)
__device__ __forceinline__ int cdf_offset(int p) { return ((p >> 2) & 1) * 128 + (p >> 3) * 4 + (p & 3); }

}  // namespace pz
