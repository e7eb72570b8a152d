// pz_model.cuh -- the linear-Gaussian model step as device functions shared by the fused step kernel
// (pz_kernels.cu) and the stored-log-weight path (pz_generic.cu): x' = F x + L z (sampleNormalDist,
// MVDistributions.hs:36-41; noisyIntegrateGen, Demo.hs:184-187) and the observation log-weight
// (mvNormal_ll0, MVDistributions.hs:45-46; gaussian_ll', Distributions.hs:160-161), every operation an
// individually rounded IEEE op in the filter's arithmetic type (DESIGN.md "Canonical math").
#pragma once
#include "pz_kernels.cuh"
#include "pz_math.cuh"

namespace pz {

// ------------------------------------------------------------------------------------------
// individually rounded IEEE operations in the filter's arithmetic type (float or double)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float r_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double r_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float r_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double r_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float r_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double r_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float r_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double r_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
// two fp32 lanes per instruction (FFMA2 / FADD2 / FMUL2, sm_100): the same individually rounded operations on each half
__device__ __forceinline__ float2 r_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 r_mul(float a, float2 b) { return __fmul2_rn(make_float2(a, a), b); }
__device__ __forceinline__ float2 r_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 r_fma(float a, float2 b, float2 c) { return __ffma2_rn(make_float2(a, a), b, c); }
__device__ __forceinline__ float2 r_sub(float a, float2 b) { return __fadd2_rn(make_float2(a, a), make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 r_sub(float2 a, float b) { return __fadd2_rn(a, make_float2(-b, -b)); }
__device__ __forceinline__ float r_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double r_max(double a, double b) { return fmax(a, b); }
template <class T> __device__ __forceinline__ const LgDevT<T>& model_of(const StepArgs& a);
template <> __device__ __forceinline__ const LgDevT<float>& model_of<float>(const StepArgs& a) { return a.model; }
template <> __device__ __forceinline__ const LgDevT<double>& model_of<double>(const StepArgs& a) { return a.modeld; }
// observation row of step t in the filter's arithmetic type
template <class T, int NY>
__device__ __forceinline__ void load_obs(const StepArgs& a, uint32_t t, T (&y)[NY]) {
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    if constexpr (sizeof(T) == 8) y[k] = a.obs64[(size_t)(t % kObsRing) * PZ_MAX_NY + k];
    else y[k] = a.obs[(size_t)(t % kObsRing) * PZ_MAX_NY + k];
  }
}

// ------------------------------------------------------------------------------------------
// linear-Gaussian model step (same operation order as the oracle's lg_propagate / lg_logweight)
// ------------------------------------------------------------------------------------------
// (C: the model's coefficient type; V: the value type -- C itself, or float2 = two slots per packed instruction)
template <int NX, class T, class V>
__device__ __forceinline__ void lg_propagate(const LgDevT<T>& p, const V* x, const V* z, V* xn) {
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    V acc = r_mul(p.F[r][0], x[0]);
#pragma unroll
    for (int c = 1; c < NX; ++c) acc = r_fma(p.F[r][c], x[c], acc);
#pragma unroll
    for (int c = 0; c <= r; ++c) acc = r_fma(p.Lq[r][c], z[c], acc);
    xn[r] = acc;
  }
}

template <int NX, int NY, class T, class V>
__device__ __forceinline__ V lg_logweight(const LgDevT<T>& p, const V* xn, const T* y) {
  V d[NY];
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    V acc = r_mul(p.H[k][0], xn[0]);
#pragma unroll
    for (int c = 1; c < NX; ++c) acc = r_fma(p.H[k][c], xn[c], acc);
    d[k] = r_sub(y[k], acc);
  }
  V quad = V();
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    V t = r_mul(p.Rinv[k][0], d[0]);
#pragma unroll
    for (int l = 1; l < NY; ++l) t = r_fma(p.Rinv[k][l], d[l], t);
    quad = (k == 0) ? r_mul(d[0], t) : r_fma(d[k], t, quad);
  }
  return r_mul(p.wscale, quad);
}

// position/velocity structured variants (LgDev::kron): same operations as the dense code minus the
// terms whose coefficient is zero
template <int NX, class T, class V>
__device__ __forceinline__ void kron_propagate(const LgDevT<T>& p, const V* x, const V* z, V* xn) {
  if constexpr (NX == 1) {  // F = 1: fl(1 * x) = x, so fma(L, z, fl(F x)) = fma(L, z, x)
    xn[0] = r_fma(p.klp, z[0], x[0]);
    return;
  }
  constexpr int P = NX / 2;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    V acc = r_mul(p.ka, x[i]);
    acc = r_fma(p.kb, x[P + i], acc);
    xn[i] = r_fma(p.klp, z[i], acc);
  }
#pragma unroll
  for (int i = 0; i < P; ++i) {
    V acc;
    if (p.kc != (T)0) { acc = r_mul(p.kc, x[i]); acc = r_fma(p.kd, x[P + i], acc); }
    else acc = r_mul(p.kd, x[P + i]);
    xn[P + i] = r_fma(p.klv, z[P + i], acc);
  }
}

template <int NX, int NY, class T, class V>
__device__ __forceinline__ V kron_logweight(const LgDevT<T>& p, const V* xn, const T* y) {
  V quad = V();
#pragma unroll
  for (int k = 0; k < NY; ++k) {
    const V d = r_sub(y[k], xn[k]);
    const V t = r_mul(p.krinv, d);
    quad = (k == 0) ? r_mul(d, t) : r_fma(d, t, quad);
  }
  return r_mul(p.wscale, quad);
}

template <int NX, bool KRON, class T, class V>
__device__ __forceinline__ void model_propagate(const LgDevT<T>& p, const V* x, const V* z, V* xn) {
  if constexpr (KRON) kron_propagate<NX>(p, x, z, xn); else lg_propagate<NX>(p, x, z, xn);
}
template <int NX, int NY, bool KRON, class T, class V>
__device__ __forceinline__ V model_logweight(const LgDevT<T>& p, const V* xn, const T* y) {
  if constexpr (KRON) return kron_logweight<NX, NY>(p, xn, y); else return lg_logweight<NX, NY>(p, xn, y);
}


// normals of ONE output slot j (the same words lg_noise4 of the fused kernel assigns to it)
template <int NX, class T>
__device__ __forceinline__ void lg_noise1(uint32_t j, uint32_t t, uint64_t seed, const float4* __restrict__ tab_m8, T (&z)[NX]) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  if constexpr (NX == 1) {
    const U4 r = philox4x32_10(j >> 2, 0u, t, kStreamProp, k0, k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    z[0] = (T)icdf_normal(w[j & 3u], tab_m8);
  } else if constexpr (NX == 6) {
    uint32_t w[12];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const U4 r = philox4x32_10(j >> 1, (uint32_t)c, t, kStreamProp, k0, k1);
      w[4 * c] = r.x; w[4 * c + 1] = r.y; w[4 * c + 2] = r.z; w[4 * c + 3] = r.w;
    }
    const int o = (j & 1u) ? 6 : 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) z[k] = (T)icdf_normal(w[o + k], tab_m8);
  } else {
#pragma unroll
    for (int c = 0; 4 * c < NX; ++c) {
      const U4 r = philox4x32_10(j, (uint32_t)c, t, kStreamProp, k0, k1);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (4 * c + i < NX) z[4 * c + i] = (T)icdf_normal(w[i], tab_m8);
    }
  }
}

}  // namespace pz
