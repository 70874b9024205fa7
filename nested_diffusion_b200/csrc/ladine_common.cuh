// Shared device helpers for the LaDiNE sampler kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/ladine.h"

namespace ladine {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kSoftplusThreshold = 20.0f;  // F.softplus(beta=1, threshold=20), latent_model.py:176

// posterior-update coefficients of one reverse step (diffusion_utils.py:69-78, :85-91), host-built
struct StepCoef {
  float inv_q, omq, s, g0, g1, g2, sig, pad;
};

// ---------------------------------------------------------------------------------------------
// softplus
// ---------------------------------------------------------------------------------------------
// exact-semantics version for the FP32 path: x > 20 ? x : log1p(exp(x))
__device__ __forceinline__ float softplus_precise(float x) {
  return x > kSoftplusThreshold ? x : log1pf(expf(x));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fast version for the tensor-core path; v2 = x * log2(e) (the caller folds log2e into scale/shift).
// 2 MUFU + 3 FP32 ops; abs error ~1e-7, far below the 16-bit operand rounding that follows.
__device__ __forceinline__ float softplus_log2dom(float v2) {
  float r = kLn2 * lg2_approx(1.0f + ex2_approx(v2));
  return v2 > kSoftplusThreshold * kLog2e ? v2 * kLn2 : r;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11), keyed on (seed; chain id, slot)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// uint32 -> uniform in (0,1): (top 24 bits + 0.5) / 2^24
__host__ __device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// standard normal number `c` (class index) of noise slot `slot` of chain `chain`.
// One Philox block yields four normals (two Box-Muller pairs); block index = c / 4.
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t chain, uint32_t slot, int c) {
  uint32_t ctr[4] = {(uint32_t)chain, (uint32_t)(chain >> 32), slot, (uint32_t)(c >> 2)};
  philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  const int pair = (c >> 1) & 1;
  const float u1 = u01(ctr[2 * pair]);
  const float u2 = u01(ctr[2 * pair + 1]);
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return (c & 1) ? r * sn : r * cs;
}

// global chain id: ((member_gid * draws_total) + draw_gid) * images_total + image_gid
struct ChainIds {
  int32_t member_gid[LADINE_MAX_GROUP];
  int32_t image_offset, images_total, draw_offset, draws_total;
  __device__ __forceinline__ uint64_t chain(int k, int d, int n) const {
    return ((uint64_t)member_gid[k] * (uint64_t)draws_total + (uint64_t)(draw_offset + d)) *
               (uint64_t)images_total + (uint64_t)(image_offset + n);
  }
};

// ---------------------------------------------------------------------------------------------
// the reverse-step update in the reference's operation order, without FMA contraction
// (diffusion_utils.py:85-91 and :105-107): every product and sum is rounded separately as the
// eager torch ops do.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float y0_reparam_op(const StepCoef& k, float y, float mu, float eps) {
  const float a = __fsub_rn(y, __fmul_rn(k.omq, mu));
  const float b = __fsub_rn(a, __fmul_rn(eps, k.s));
  return __fmul_rn(k.inv_q, b);
}
__device__ __forceinline__ float posterior_step_op(const StepCoef& k, float y, float mu, float eps, float z) {
  const float y0r = y0_reparam_op(k, y, mu, eps);
  const float m = __fadd_rn(__fadd_rn(__fmul_rn(k.g0, y0r), __fmul_rn(k.g1, y)), __fmul_rn(k.g2, mu));
  return __fadd_rn(m, __fmul_rn(k.sig, z));
}

// 16-bit operand packing for the tensor-core path.  Conversions SATURATE (cvt.rn.satfinite: one F2FP instruction,
// |v| > max finite -> +-max finite, 65504 for FP16) instead of rounding to infinity: an overflowing activation would
// otherwise reach the next tcgen05 GEMM as inf and come out as NaN (inf x mixed-sign weights).  NaN stays NaN.
template <typename T>
struct Pack16;
template <>
struct Pack16<__half> {
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ __half one(float v) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_half(r);
  }
};
template <>
struct Pack16<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ __nv_bfloat16 one(float v) {
    unsigned short r;
    asm("cvt.rn.satfinite.bf16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_bfloat16(r);
  }
};

}  // namespace ladine
