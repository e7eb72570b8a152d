// pz_math.cuh -- device side of the canonical math spec (DESIGN.md "Math" and "RNG").
//
// Counter-based Philox4x32-10 replaces the reference's sequential MWC256 SamplerIO
// (monad-bayes Control.Monad.Bayes.Sampler; call sites Inference.hs:30-42): a draw is a pure
// function of (seed, global slot, step, call), so particles are reproducible independently
// of launch shape and GPU count.
//
// Every fp32 operation here is an individually rounded IEEE op (the translation unit is
// compiled with -fmad=false); fused multiply-adds appear exactly where __fmaf_rn is
// written.  The CPU oracle carries the same literals and the same operation order, which
// is what makes whole-filter trajectories bit-comparable.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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
constexpr float kNeg2Ln2 = -0x1.62e430p+0f;
constexpr float kHalfPi = 0x1.921fb6p+0f;

constexpr int32_t kNZero = -(1 << 30);  // exponent marker of a zero weight

// log-weight -> (n, m): weight = m * 2^(n-23), m in [2^22.5, 2^23.5].
// -inf / +inf / NaN / out-of-range -> (kNZero, 0).
__device__ __forceinline__ void weight_split(float lw, int32_t& n, uint32_t& m) {
  const float e = __fmul_rn(lw, kLog2e);
  const float r = rintf(e);
  const float f = __fsub_rn(e, r);
  float p = 0x1.00c0e6p-16f;
  p = __fmaf_rn(p, f, 0x1.444004p-13f);
  p = __fmaf_rn(p, f, 0x1.5d8776p-10f);
  p = __fmaf_rn(p, f, 0x1.3b2a1cp-7f);
  p = __fmaf_rn(p, f, 0x1.c6b08ep-5f);
  p = __fmaf_rn(p, f, 0x1.ebfbe0p-3f);
  p = __fmaf_rn(p, f, 0x1.62e430p-1f);
  p = __fmaf_rn(p, f, 1.0f);
  const bool ok = fabsf(e) < 1.0e6f;  // false for NaN and +-inf
  n = ok ? __float2int_rn(r) : kNZero;
  m = ok ? (uint32_t)__float2int_rn(__fmul_rn(p, 8388608.0f)) : 0u;
}

// fixed-point weight of (n, m) at block exponent eb >= n: m >> (eb - n), 0 once the shift
// reaches 32 (PTX shr clamps the shift amount, so no compare is needed)
__device__ __forceinline__ uint32_t q_at(uint32_t m, int32_t n, int32_t eb) {
  uint32_t q;
  asm("shr.u32 %0, %1, %2;" : "=r"(q) : "r"(m), "r"((uint32_t)(eb - n)));
  return q;
}

// two 32-bit words -> two independent N(0,1) draws
__device__ __forceinline__ void normal_pair(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  float u = __fmaf_rn(__uint2float_rn(ra), 0x1p-32f, 0x1p-33f);
  uint32_t bits = __float_as_uint(u);
  int32_t e = (int32_t)(bits - 0x3F3504F3u) >> 23;
  float mant = __uint_as_float(bits - ((uint32_t)e << 23));
  float t = __fsub_rn(mant, 1.0f);
  float t2 = __fmul_rn(t, t);
  float P = -0x1.4bd700p-4f;
  P = __fmaf_rn(P, t, 0x1.04591ap-3f);
  P = __fmaf_rn(P, t, -0x1.09abe4p-3f);
  P = __fmaf_rn(P, t, 0x1.22dc4cp-3f);
  P = __fmaf_rn(P, t, -0x1.54d54ep-3f);
  P = __fmaf_rn(P, t, 0x1.99a14cp-3f);
  P = __fmaf_rn(P, t, -0x1.0000c8p-2f);
  P = __fmaf_rn(P, t, 0x1.555552p-2f);
  float lnm = __fmaf_rn(__fmul_rn(t2, t), P, __fmaf_rn(-0.5f, t2, t));
  float rad2 = __fmaf_rn(__int2float_rn(e), kNeg2Ln2, __fmul_rn(-2.0f, lnm));
  rad2 = fmaxf(rad2, 0.0f);
  float R = __fsqrt_rn(rad2);
  uint32_t q = rb >> 30;
  float fr = __fmaf_rn(__uint2float_rn(rb & 0x3FFFFFFFu), 0x1p-30f, -0.5f);
  float phi = __fmul_rn(fr, kHalfPi);
  float z = __fmul_rn(phi, phi);
  float S = __fmaf_rn(__fmaf_rn(-0x1.9aca02p-13f, z, 0x1.110c2ap-7f), z, -0x1.555552p-3f);
  float C = __fmaf_rn(__fmaf_rn(0x1.9bd8d2p-16f, z, -0x1.6c12d4p-10f), z, 0x1.555554p-5f);
  float s = __fmaf_rn(__fmul_rn(phi, z), S, phi);
  float c = __fmaf_rn(__fmul_rn(z, z), C, __fmaf_rn(-0.5f, z, 1.0f));
  float cq = (q & 1u) ? s : c;   // q=0: c   1: -s   2: -c   3: s
  float sq = (q & 1u) ? c : s;   // q=0: s   1:  c   2: -s   3: -c
  if (q == 1u || q == 2u) cq = -cq;
  if (q >= 2u) sq = -sq;
  z0 = __fmul_rn(R, cq);
  z1 = __fmul_rn(R, sq);
}

}  // namespace pz
