// Shared GEMM epilogue (used by the SIMT and the tcgen05 GEMM kernels).
#pragma once
#include "common.cuh"

namespace anyloc {

struct EpiParams {
  int mode;
  const float* bias;    // [N] nullable
  const float* gamma;   // [N] (LS_RESID)
  const float* resid;   // [M,ldo] (LS_RESID; may alias out)
  float* out;           // [M,ldo]
  float* out_lo;        // [M,ldo] (SPLIT modes)
  int ldo;
  // QKV_SPLIT only: columns >= 2*qkv_D (the V third) are written TRANSPOSED per head into
  // vt_{hi,lo}[(b*heads + h)*64 + d][t] (row pitch qkv_Tp) so the attention kernel can use V as a K-major operand
  float* vt_hi = nullptr;
  float* vt_lo = nullptr;
  int qkv_T = 0, qkv_Tp = 0, qkv_D = 0;
  int qkv_f16 = 0;      // QKV_SPLIT: q,k and V^T as fp16 pairs of kActScale*x (out/out_lo/vt_* point to __half)
  float alpha = 1.0f;   // accumulator scale (1/(s_A*s_B) for fp16-pair inputs, else 1)
  int out_f16 = 0;      // SPLIT outputs as fp16 pairs of kActScale*v (out/out_lo then point to __half)
  const int* gate = nullptr;   // device flag (nullable): tcgen05 kernels return immediately when *gate == 0 (conditional fallbacks without a host sync)
};

__device__ __forceinline__ void epi_store_split(const EpiParams& p, size_t o, float v) {
  if (p.out_f16) {
    __half h, l; split_f16(v * kActScale, h, l);
    reinterpret_cast<__half*>(p.out)[o] = h; reinterpret_cast<__half*>(p.out_lo)[o] = l;
  } else {
    float h, l; split_tf32(v, h, l); p.out[o] = h; p.out_lo[o] = l;
  }
}

__device__ __forceinline__ void epi_store_vt(const EpiParams& p, int m, int n, float v) {
  // n in [2D, 3D): head h = (n-2D)/64, dim d = (n-2D)%64; row m = b*T + t
  const int c = n - 2 * p.qkv_D, b = m / p.qkv_T, t = m - b * p.qkv_T;
  const size_t o = ((size_t)b * p.qkv_D + c) * p.qkv_Tp + t;      // (b*heads + h)*64 + d == b*D + c
  if (p.qkv_f16) {
    __half h, l; split_f16(v * kActScale, h, l);
    reinterpret_cast<__half*>(p.vt_hi)[o] = h; reinterpret_cast<__half*>(p.vt_lo)[o] = l;
  } else {
    float h, l; split_tf32(v, h, l);
    p.vt_hi[o] = h; p.vt_lo[o] = l;
  }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
// Same function on the special-function unit: e = 2^(-x log2 e) (ex2.approx, 2^-22 relative), 1/(1+e) by rcp.approx +
// one Newton step.  |error| <= ~3e-7 |silu(x)|; 7 instructions instead of ~35 (the SwiGLU epilogue runs it 4 M times
// per GEMM).  x -> -inf: e = inf, 1/(1+e) = 0 -> -0;  NaN propagates.
__device__ __forceinline__ float silu_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  const float d = 1.0f + e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = (d < 3.0e38f) ? fmaf(r, fmaf(-d, r, 1.0f), r) : r;     // Newton step (skipped when d overflowed: r = 0)
  return x * r;
}

// Apply the epilogue to one accumulator element (m, n).  For SWIGLU the caller passes the PAIR
// (acc0 at column n even, acc1 at column n+1) and the result lands in column n/2.
__device__ __forceinline__ void epi_store1(const EpiParams& p, int m, int n, float acc) {
  float v = acc * p.alpha + (p.bias ? __ldg(p.bias + n) : 0.f);
  size_t o = (size_t)m * p.ldo + n;
  switch (p.mode) {
    case ANYLOC_EPI_BIAS: p.out[o] = v; break;
    case ANYLOC_EPI_BIAS_SPLIT: epi_store_split(p, o, v); break;
    case ANYLOC_EPI_GELU_SPLIT: epi_store_split(p, o, gelu_erf(v)); break;
    case ANYLOC_EPI_LS_RESID: p.out[o] = p.resid[o] + __ldg(p.gamma + n) * v; break;
    case ANYLOC_EPI_QKV_SPLIT:
      if (n >= 2 * p.qkv_D) epi_store_vt(p, m, n, v);
      else if (p.qkv_f16) {
        __half h, l; split_f16(v * kActScale, h, l);
        reinterpret_cast<__half*>(p.out)[o] = h; reinterpret_cast<__half*>(p.out_lo)[o] = l;
      } else { float h, l; split_tf32(v, h, l); p.out[o] = h; p.out_lo[o] = l; }
      break;
    default: break;
  }
}
__device__ __forceinline__ void epi_store_pair(const EpiParams& p, int m, int n_even, float acc0, float acc1) {
  // SWIGLU: columns (n_even, n_even+1) = (x1_j, x2_j), j = n_even/2
  float x1 = acc0 * p.alpha + (p.bias ? __ldg(p.bias + n_even) : 0.f);
  float x2 = acc1 * p.alpha + (p.bias ? __ldg(p.bias + n_even + 1) : 0.f);
  epi_store_split(p, (size_t)m * p.ldo + (n_even >> 1), silu(x1) * x2);
}

}  // namespace anyloc
