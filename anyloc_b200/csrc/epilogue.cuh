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
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// Apply the epilogue to one accumulator element (m, n).  For SWIGLU the caller passes the PAIR
// (acc0 at column n even, acc1 at column n+1) and the result lands in column n/2.
__device__ __forceinline__ void epi_store1(const EpiParams& p, int m, int n, float acc) {
  float v = acc + (p.bias ? __ldg(p.bias + n) : 0.f);
  size_t o = (size_t)m * p.ldo + n;
  switch (p.mode) {
    case ANYLOC_EPI_BIAS: p.out[o] = v; break;
    case ANYLOC_EPI_BIAS_SPLIT: { float h, l; split_tf32(v, h, l); p.out[o] = h; p.out_lo[o] = l; } break;
    case ANYLOC_EPI_GELU_SPLIT: { float h, l; split_tf32(gelu_erf(v), h, l); p.out[o] = h; p.out_lo[o] = l; } break;
    case ANYLOC_EPI_LS_RESID: p.out[o] = p.resid[o] + __ldg(p.gamma + n) * v; break;
    default: break;
  }
}
__device__ __forceinline__ void epi_store_pair(const EpiParams& p, int m, int n_even, float acc0, float acc1) {
  // SWIGLU: columns (n_even, n_even+1) = (x1_j, x2_j), j = n_even/2
  float x1 = acc0 + (p.bias ? __ldg(p.bias + n_even) : 0.f);
  float x2 = acc1 + (p.bias ? __ldg(p.bias + n_even + 1) : 0.f);
  float v = silu(x1) * x2, h, l;
  split_tf32(v, h, l);
  size_t o = (size_t)m * p.ldo + (n_even >> 1);
  p.out[o] = h; p.out_lo[o] = l;
}

}  // namespace anyloc
