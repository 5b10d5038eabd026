// Query->database retrieval (reference: /root/reference/utilities.py:435-450, faiss IndexFlatIP /
// IndexFlatL2 exact search).  normalise rows -> score GEMM (fp32-equivalent) -> k-best per query,
// best first, lowest database index first among equal scores.
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"

// internal (api.cu): the plain-store 3-term GEMM with an optional device gate (*gate == 0 -> the kernels return at once);
// when gated, no algorithmic work is recorded (the coarse pass already accounted for the product)
extern "C" int anyloc_gemm_nt_gated(const void* a_hi, const void* a_lo, int lda, const void* b_hi, const void* b_lo, int ldb,
                                    int M, int N, int K, int in_dtype, float alpha, float* out, int ldo, const int* gate,
                                    void* stream);

namespace anyloc {

// fp16-pair scale of unit-norm rows: |s y| <= 4096 < 65504, and s*y - hi stays far above the fp16 subnormal step for
// every element that matters (an element below 2^-12 of the row norm contributes < 2^-36 to a unit dot product).
constexpr float kRetrievalScale = 4096.0f;

// y = x / max(|x|, 1e-12) written as a tf32 (hi,lo) pair (hi+lo == fp32 value); also |y|^2 per row
// (needed by the L2 metric).  One CTA per row.
template <bool F16>
__global__ void __launch_bounds__(256)
normalize_rows_split_kernel(const float* __restrict__ x, int D, int do_norm, void* __restrict__ hi_v,
                            void* __restrict__ lo_v, float* __restrict__ sq, float* __restrict__ dn /* nullable, F16 only */,
                            int* __restrict__ dn_max_bits /* nullable */, int n_rows) {
  // grid-stride over the rows with a grid of ~3 CTAs per SM: a row (196 KB at Dv = 49152) is read twice (norm pass,
  // split pass) and the second read must still find it in L2 -- with one CTA per row ~1200 rows (230 MB) were in
  // flight and both passes went to DRAM (ncu, round 2: 3.93 GB read for a 1.97 GB database)
  for (size_t row = blockIdx.x; row < (size_t)n_rows; row += gridDim.x) {
  __syncthreads();
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  const int D4 = D >> 2;
  float ss = 0.f;
  for (int d = threadIdx.x; d < D4; d += blockDim.x) {
    float4 v = __ldg(xr + d);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  __shared__ float red[8];
  __shared__ float tot;
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; tot = t; }
  __syncthreads();
  const float nrm = fmaxf(sqrtf(tot), 1e-12f);
  float4* h4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(hi_v) + row * D);
  float4* l4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(lo_v) + row * D);
  uint2* h2 = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi_v) + row * D);   // 4 halves = 8 bytes
  uint2* l2 = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo_v) + row * D);
  float ss2 = 0.f, dd2 = 0.f;
  for (int d = threadIdx.x; d < D4; d += blockDim.x) {
    float4 v = __ldg(xr + d);
    if (do_norm) { v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm; }
    ss2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if constexpr (F16) {
      uint2 h, l;
      const float a0 = v.x * kRetrievalScale, a1 = v.y * kRetrievalScale, a2 = v.z * kRetrievalScale, a3 = v.w * kRetrievalScale;
      split_f16x2(a0, a1, h.x, l.x);
      split_f16x2(a2, a3, h.y, l.y);
      h2[d] = h; l2[d] = l;
      // what a hi-only product drops of this row: s*y - hi (exact in fp32: hi is s*y rounded to 11 significant bits)
      const float e0 = a0 - veltkamp_hi11(a0), e1 = a1 - veltkamp_hi11(a1), e2 = a2 - veltkamp_hi11(a2), e3 = a3 - veltkamp_hi11(a3);
      dd2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    } else {
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y);
      split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      h4[d] = h; l4[d] = l;
    }
  }
  if (sq) {
    ss2 = warp_sum(ss2);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss2;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; sq[row] = t; }
  }
  if (F16 && dn) {
    // |s*y - hi| / s, rounded UP (1 + 2^-10), plus the fp16 subnormal slack sqrt(D) * 2^-25 / s (elements of |s*y| < 2^-14
    // are rounded to the 2^-24 grid) -- a rigorous bound on what the hi-only (coarse) score ignores of this row
    dd2 = warp_sum(dd2);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dd2;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w];
      const float b = (sqrtf(t) * 1.001f + sqrtf((float)D) * 2.98e-8f) / kRetrievalScale;
      dn[row] = b;
      if (dn_max_bits) atomicMax(dn_max_bits, __float_as_int(b));      // positive floats order like their bit patterns
    }
  }
  }
}

// Single-pass form for rows that fit the registers of a 1024-thread CTA (D <= 16 * 4096 floats): the row is loaded ONCE
// (NV float4 per thread, the whole row in flight), reduced, normalised, split and stored -- 4 B read + 4 B (fp16 pairs) or
// 8 B (tf32 pairs) written per element, instead of reading the row twice.  Same arithmetic, except that |x|^2 is summed in
// the order of this thread layout (fp32 rounding of the norm only).
template <bool F16, int NV>
__global__ void __launch_bounds__(1024, 1)
normalize_rows_split_reg_kernel(const float* __restrict__ x, int D, int do_norm, void* __restrict__ hi_v,
                                void* __restrict__ lo_v, float* __restrict__ sq, float* __restrict__ dn,
                                int* __restrict__ dn_max_bits, int n_rows) {
  const int D4 = D >> 2;
  __shared__ float red[32];
  __shared__ float s_tot;
  for (size_t row = blockIdx.x; row < (size_t)n_rows; row += gridDim.x) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    float4 v[NV];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int d = threadIdx.x + j * 1024;
      v[j] = d < D4 ? __ldg(xr + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    ss = warp_sum(ss);
    __syncthreads();                                   // red / s_tot of the previous row have been consumed
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 32; ++w) t += red[w]; s_tot = t; }
    __syncthreads();
    const float nrm = fmaxf(sqrtf(s_tot), 1e-12f);
    float ss2 = 0.f, dd2 = 0.f;
    uint2* h2 = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi_v) + row * D);
    uint2* l2 = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo_v) + row * D);
    float4* h4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(hi_v) + row * D);
    float4* l4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(lo_v) + row * D);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int d = threadIdx.x + j * 1024;
      if (d >= D4) continue;
      float4 y = v[j];
      if (do_norm) { y.x = y.x / nrm; y.y = y.y / nrm; y.z = y.z / nrm; y.w = y.w / nrm; }
      ss2 += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
      if constexpr (F16) {
        uint2 h, l;
        const float a0 = y.x * kRetrievalScale, a1 = y.y * kRetrievalScale, a2 = y.z * kRetrievalScale, a3 = y.w * kRetrievalScale;
        split_f16x2(a0, a1, h.x, l.x);
        split_f16x2(a2, a3, h.y, l.y);
        h2[d] = h; l2[d] = l;
        const float e0 = a0 - veltkamp_hi11(a0), e1 = a1 - veltkamp_hi11(a1), e2 = a2 - veltkamp_hi11(a2), e3 = a3 - veltkamp_hi11(a3);
        dd2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      } else {
        float4 h, l;
        split_tf32(y.x, h.x, l.x); split_tf32(y.y, h.y, l.y); split_tf32(y.z, h.z, l.z); split_tf32(y.w, h.w, l.w);
        h4[d] = h; l4[d] = l;
      }
    }
    ss2 = warp_sum(ss2); dd2 = warp_sum(dd2);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss2;
    __syncthreads();
    if (threadIdx.x == 0 && sq) { float t = 0.f; for (int w = 0; w < 32; ++w) t += red[w]; sq[row] = t; }
    if (F16 && dn) {
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dd2;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f; for (int w = 0; w < 32; ++w) t += red[w];
        const float b = (sqrtf(t) * 1.001f + sqrtf((float)D) * 2.98e-8f) / kRetrievalScale;
        dn[row] = b;
        if (dn_max_bits) atomicMax(dn_max_bits, __float_as_int(b));
      }
    }
  }
}

// host dispatch of the two forms
template <bool F16>
static cudaError_t launch_normalize_rows(const float* x, int n_rows, int D, int do_norm, void* hi, void* lo, float* sq, float* dn,
                                         int* dn_max_bits, cudaStream_t st) {
  const int D4 = D >> 2, sms = device_sm_count();
  if (D4 > 2 * 1024 && D4 <= 16 * 1024) {            // long rows: the single-pass register form, one CTA per SM
    const int grid = std::min(n_rows, sms);
    if (D4 <= 4 * 1024) normalize_rows_split_reg_kernel<F16, 4><<<grid, 1024, 0, st>>>(x, D, do_norm, hi, lo, sq, dn, dn_max_bits, n_rows);
    else if (D4 <= 8 * 1024) normalize_rows_split_reg_kernel<F16, 8><<<grid, 1024, 0, st>>>(x, D, do_norm, hi, lo, sq, dn, dn_max_bits, n_rows);
    else if (D4 <= 12 * 1024) normalize_rows_split_reg_kernel<F16, 12><<<grid, 1024, 0, st>>>(x, D, do_norm, hi, lo, sq, dn, dn_max_bits, n_rows);
    else normalize_rows_split_reg_kernel<F16, 16><<<grid, 1024, 0, st>>>(x, D, do_norm, hi, lo, sq, dn, dn_max_bits, n_rows);
  } else {
    normalize_rows_split_kernel<F16><<<std::min(n_rows, 8 * sms), 256, 0, st>>>(x, D, do_norm, hi, lo, sq, dn, dn_max_bits, n_rows);
  }
  return cudaGetLastError();
}

// k-best selection per query row, best first, lowest index first among equal scores (metric IP: larger is better;
// metric L2: key := -(qq - 2 s + dd)).  O(n_db) per query:
//   A: every thread keeps the best score of its strided share of the row;
//   B: tau = the k-th best of the 1024 thread bests (k rounds of block arg-best over 1024 values): those are k DISTINCT
//      row elements, so the row's true k-th best is >= tau and every member of the true top-k is >= tau;
//   C: second sweep, the elements >= tau (usually k .. a few k of them) are compacted into a shared-memory list;
//   D: k rounds of arg-best over the list (score descending, lowest index first among equal scores).
// More than SEL_CAP candidates (a row with thousands of equal scores) -> k ordered sweeps over the row (read-only).
constexpr int SEL_CAP = 4096;
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void block_argbest(float& best, int& besti, float* bv, int* bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (better(ov, oi, best, besti)) { best = ov; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? bv[threadIdx.x] : -INFINITY;
    besti = threadIdx.x < nw ? bi[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (better(ov, oi, best, besti)) { best = ov; besti = oi; }
    }
    if (threadIdx.x == 0) { bv[32] = best; bi[32] = besti; }
  }
  __syncthreads();
  best = bv[32]; besti = bi[32];
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
topk_select2_kernel(const float* __restrict__ scores, int n_db, int64_t ld, int k, int metric,
                    const float* __restrict__ qq, const float* __restrict__ dd,
                    float* __restrict__ dist, int64_t* __restrict__ idx, const int* __restrict__ gate /* nullable */) {
  if (gate != nullptr && *reinterpret_cast<const volatile int*>(gate) == 0) return;
  const int q = blockIdx.x;
  const float* s = scores + (size_t)q * ld;
  const bool l2 = metric == ANYLOC_METRIC_L2;
  const float a = l2 ? qq[q] : 0.f;
  auto key = [&](int j) { const float v = s[j]; return l2 ? -((a - 2.0f * v) + dd[j]) : v; };   // larger = better
  __shared__ float bv[33];
  __shared__ int bi[33];
  __shared__ float cv[SEL_CAP];
  __shared__ int ci[SEL_CAP];
  __shared__ int count;
  if (threadIdx.x == 0) count = 0;
  // A
  float mine = -INFINITY; int minei = 0x7fffffff;
  for (int j = threadIdx.x; j < n_db; j += blockDim.x) {
    const float v = key(j);
    if (better(v, j, mine, minei)) { mine = v; minei = j; }
  }
  // B
  float tau = -INFINITY;
  {
    float v = mine; int vi = minei;
    for (int r = 0; r < k; ++r) {
      float b = v; int bidx = vi;
      block_argbest(b, bidx, bv, bi);
      if (bidx == 0x7fffffff) { tau = -INFINITY; break; }   // fewer than k elements in the row
      tau = b;
      if (vi == bidx) { v = -INFINITY; vi = 0x7fffffff; }    // the winner leaves the pool
    }
  }
  __syncthreads();
  // C
  for (int j = threadIdx.x; j < n_db; j += blockDim.x) {
    const float v = key(j);
    if (v >= tau) {
      const int slot = atomicAdd(&count, 1);
      if (slot < SEL_CAP) { cv[slot] = v; ci[slot] = j; }
    }
  }
  __syncthreads();
  const int n_c = count;
  if (n_c > SEL_CAP) {
    // thousands of (near-)equal scores: k ordered sweeps over the row itself -- round r takes the best element that
    // comes strictly after round r-1's winner in (score descending, index ascending) order
    float pv = INFINITY; int pj = -1;
    for (int r = 0; r < k; ++r) {
      float b = -INFINITY; int bidx = 0x7fffffff;
      for (int j = threadIdx.x; j < n_db; j += blockDim.x) {
        const float v = key(j);
        if ((v < pv || (v == pv && j > pj)) && better(v, j, b, bidx)) { b = v; bidx = j; }
      }
      block_argbest(b, bidx, bv, bi);
      if (threadIdx.x == 0) {
        dist[(size_t)q * k + r] = bidx != 0x7fffffff ? (l2 ? -b : b) : (l2 ? INFINITY : -INFINITY);
        idx[(size_t)q * k + r] = bidx != 0x7fffffff ? bidx : -1;
      }
      pv = b; pj = bidx;
      if (bidx == 0x7fffffff) pv = -INFINITY;      // exhausted: every later round pads
    }
    return;
  }
  // D
  for (int r = 0; r < k; ++r) {
    float b = -INFINITY; int bidx = 0x7fffffff; int bslot = -1;
    for (int c = threadIdx.x; c < n_c; c += blockDim.x)
      if (better(cv[c], ci[c], b, bidx)) { b = cv[c]; bidx = ci[c]; bslot = c; }
    float wb = b; int wi = bidx;
    block_argbest(wb, wi, bv, bi);
    if (bslot >= 0 && wi == bidx && wi != 0x7fffffff) { cv[bslot] = -INFINITY; ci[bslot] = 0x7fffffff; }
    if (threadIdx.x == 0) {
      if (wi != 0x7fffffff) {
        dist[(size_t)q * k + r] = l2 ? -wb : wb;
        idx[(size_t)q * k + r] = wi;
      } else {
        dist[(size_t)q * k + r] = l2 ? INFINITY : -INFINITY;
        idx[(size_t)q * k + r] = -1;          // faiss pads with -1 when k > ntotal
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Coarse retrieval (inner product on an fp16-pair index).  One hi-only tensor-core pass gives S~ = hi_q . hi_d / s^2 with
//   |S~ - S| <= eps_q = (dn_q + DN + dn_q DN) (1 + 2^-10) + 3e-5
// (dn_* = |s y - hi| / s per row, measured when the pairs were written; DN = max over the database; Cauchy-Schwarz per
// dropped term, |y| <= 1 + 1e-6; 3e-5 covers the fp32 accumulation: <= 128 truncating steps per TMEM chunk + 24
// round-to-nearest chunk adds).  With tau = the k-th best S~ of a query, every member of its exact top-k has
// S~ >= tau - 2 eps_q, so those candidates (usually k .. 3k of 10^4-10^5 rows) are re-scored EXACTLY from the (hi,lo)
// pairs in fp32 and the k best of them -- score descending, lowest index first -- are the answer: identical to the
// 3-term path up to fp32 rounding of the scores, at a third of the tensor-core work.  A query with more than CAND_MAX
// candidates raises a device flag that switches on the 3-term fallback (launched behind it, gated, no host sync).
constexpr int CAND_MAX = 256;

__global__ void __launch_bounds__(1024)
topk_candidates_kernel(const float* __restrict__ scores, int n_db, int64_t ld, int k, const float* __restrict__ dn_q,
                       const int* __restrict__ dn_max_bits, int32_t* __restrict__ cand /* [n_q, CAND_MAX] */,
                       int32_t* __restrict__ cand_n /* [n_q] */, int* __restrict__ overflow) {
  const int q = blockIdx.x;
  const float* s = scores + (size_t)q * ld;
  __shared__ float bv[33];
  __shared__ int bi[33];
  __shared__ int ci[SEL_CAP];
  __shared__ int count;
  if (threadIdx.x == 0) count = 0;
  float mine = -INFINITY; int minei = 0x7fffffff;
  for (int j = threadIdx.x; j < n_db; j += blockDim.x) {
    const float v = s[j];
    if (better(v, j, mine, minei)) { mine = v; minei = j; }
  }
  float tau = -INFINITY;
  {
    float v = mine; int vi = minei;
    for (int r = 0; r < k; ++r) {
      float b = v; int bidx = vi;
      block_argbest(b, bidx, bv, bi);
      if (bidx == 0x7fffffff) { tau = -INFINITY; break; }
      tau = b;
      if (vi == bidx) { v = -INFINITY; vi = 0x7fffffff; }
    }
  }
  const float DN = __int_as_float(*dn_max_bits), dq = dn_q[q];
  const float eps = (dq + DN + dq * DN) * 1.001f + 3.0e-5f;
  const float thr = tau - 2.0f * eps;
  __syncthreads();
  for (int j = threadIdx.x; j < n_db; j += blockDim.x) {
    if (s[j] >= thr) {
      const int slot = atomicAdd(&count, 1);
      if (slot < SEL_CAP) ci[slot] = j;
    }
  }
  __syncthreads();
  const int n_c = count;
  if (n_c > CAND_MAX) {
    if (threadIdx.x == 0) { cand_n[q] = 0; atomicExch(overflow, 1); }
    return;
  }
  for (int c = threadIdx.x; c < n_c; c += blockDim.x) cand[(size_t)q * CAND_MAX + c] = ci[c];
  if (threadIdx.x == 0) cand_n[q] = n_c;
}

// CTA per query: exact fp32 scores of its candidates from the fp16 (hi,lo) pairs, then the k best of them.
__device__ __forceinline__ float2 h2f(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
constexpr int RS_CHUNK = 4096;          // query elements staged per pass (fp32: 16 KB of shared memory)
__global__ void __launch_bounds__(256)
topk_rescore_kernel(const __half* __restrict__ db_hi, const __half* __restrict__ db_lo, const __half* __restrict__ qu_hi,
                    const __half* __restrict__ qu_lo, int Dv, int k, const int32_t* __restrict__ cand,
                    const int32_t* __restrict__ cand_n, const int* __restrict__ overflow, float* __restrict__ dist,
                    int64_t* __restrict__ idx) {
  if (*reinterpret_cast<const volatile int*>(overflow) != 0) return;       // the 3-term fallback answers every query
  const int q = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ __align__(16) float xq[RS_CHUNK];      // this pass' slice of the query, hi + lo (exact: 22 significant bits)
  __shared__ float cv[CAND_MAX];
  __shared__ int ci[CAND_MAX];
  __shared__ float bv[33];
  __shared__ int bi[33];
  const int n_c = cand_n[q];
  for (int c = threadIdx.x; c < n_c; c += blockDim.x) { cv[c] = 0.f; ci[c] = cand[(size_t)q * CAND_MAX + c]; }
  // The query is read ONCE per CTA (slice by slice into shared memory); every candidate row is streamed once, in
  // 8 KB pieces per array.  Warp w owns candidates w, w + 8, ...; partial sums are added slice by slice in slice order.
  for (int d0 = 0; d0 < Dv; d0 += RS_CHUNK) {
    const int len = min(RS_CHUNK, Dv - d0);          // multiple of 8
    __syncthreads();
    for (int t = threadIdx.x; t < (len >> 3); t += blockDim.x) {
      const uint4 xh = __ldg(reinterpret_cast<const uint4*>(qu_hi + (size_t)q * Dv + d0) + t);
      const uint4 xl = __ldg(reinterpret_cast<const uint4*>(qu_lo + (size_t)q * Dv + d0) + t);
      const uint32_t hw[4] = {xh.x, xh.y, xh.z, xh.w}, lw[4] = {xl.x, xl.y, xl.z, xl.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 a = h2f(hw[e]), b2 = h2f(lw[e]); o[2 * e] = a.x + b2.x; o[2 * e + 1] = a.y + b2.y; }
      *reinterpret_cast<float4*>(xq + t * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(xq + t * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    for (int c = w; c < n_c; c += 8) {
      const size_t roff = (size_t)ci[c] * Dv + d0;
      const uint4* dh = reinterpret_cast<const uint4*>(db_hi + roff);
      const uint4* dl = reinterpret_cast<const uint4*>(db_lo + roff);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int t = lane; t < (len >> 3); t += 32) {
        const uint4 yh = __ldg(dh + t), yl = __ldg(dl + t);
        const float4 x0 = *reinterpret_cast<const float4*>(xq + t * 8), x1 = *reinterpret_cast<const float4*>(xq + t * 8 + 4);
        const float2 h0 = h2f(yh.x), l0 = h2f(yl.x), h1 = h2f(yh.y), l1 = h2f(yl.y);
        const float2 h2 = h2f(yh.z), l2 = h2f(yl.z), h3 = h2f(yh.w), l3 = h2f(yl.w);
        a0 = fmaf(x0.x, h0.x + l0.x, a0); a1 = fmaf(x0.y, h0.y + l0.y, a1);
        a2 = fmaf(x0.z, h1.x + l1.x, a2); a3 = fmaf(x0.w, h1.y + l1.y, a3);
        a0 = fmaf(x1.x, h2.x + l2.x, a0); a1 = fmaf(x1.y, h2.y + l2.y, a1);
        a2 = fmaf(x1.z, h3.x + l3.x, a2); a3 = fmaf(x1.w, h3.y + l3.y, a3);
      }
      const float part = warp_sum((a0 + a1) + (a2 + a3));
      if (lane == 0) cv[c] += part;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_c; c += blockDim.x) cv[c] *= 1.0f / (kRetrievalScale * kRetrievalScale);
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float b = -INFINITY; int bidx = 0x7fffffff; int bslot = -1;
    for (int c = threadIdx.x; c < n_c; c += blockDim.x)
      if (better(cv[c], ci[c], b, bidx)) { b = cv[c]; bidx = ci[c]; bslot = c; }
    float wb = b; int wi = bidx;
    block_argbest(wb, wi, bv, bi);
    if (bslot >= 0 && wi == bidx && wi != 0x7fffffff) { cv[bslot] = -INFINITY; ci[bslot] = 0x7fffffff; }
    if (threadIdx.x == 0) {
      dist[(size_t)q * k + r] = wi != 0x7fffffff ? wb : -INFINITY;
      idx[(size_t)q * k + r] = wi != 0x7fffffff ? wi : -1;
    }
    __syncthreads();
  }
}

}  // namespace anyloc

using namespace anyloc;


// ---- prepared database ("index"): what faiss' index.add(db) leaves behind (utilities.py:449).
// Layout of the caller-owned blob for `capacity` rows of Dv columns:
//   hi [capacity, Dv] | lo [capacity, Dv] | sq [capacity] fp32 (|y|^2 per row, the L2 metric needs it) |
//   dn [capacity] fp32 (|s y - hi| / s per row: what a hi-only product ignores; fp16 layout only) | header (256 B: max dn)
// with hi/lo either fp16 pairs of 4096*y (unit rows: normalize != 0, Dv % 8 == 0 -- the 2x faster kind::f16 tensor
// path) or tf32 pairs of y (rows of unknown range).  ANYLOC_TOPK_F16=0 forces the tf32 pairs (A/B).
static bool index_uses_f16(int Dv, int normalize) {
  static int f16_env = -1;
  if (f16_env < 0) { const char* e = getenv("ANYLOC_TOPK_F16"); f16_env = e ? atoi(e) : 1; }
  return f16_env && normalize && (Dv % 8) == 0;
}
struct IndexView { void* hi; void* lo; float* sq; float* dn; int* hdr; bool f16; size_t pair_bytes_per_row; };
static bool carve_index(void* blob, size_t bytes, int64_t capacity, int Dv, int normalize, IndexView* v) {
  v->f16 = index_uses_f16(Dv, normalize);
  const size_t esz = v->f16 ? 2 : 4;
  v->pair_bytes_per_row = (size_t)Dv * esz;
  Workspace w(blob, bytes);
  v->hi = w.take<char>((size_t)capacity * Dv * esz);
  v->lo = w.take<char>((size_t)capacity * Dv * esz);
  v->sq = w.take<float>((size_t)capacity);
  v->dn = w.take<float>((size_t)capacity);
  v->hdr = w.take<int>(64);
  return v->hi && v->lo && v->sq && v->dn && v->hdr;
}

extern "C" size_t anyloc_index_bytes(int64_t capacity, int Dv, int normalize) {
  const size_t esz = index_uses_f16(Dv, normalize) ? 2 : 4;
  return 2 * align_up((size_t)capacity * Dv * esz, 256) + 2 * align_up((size_t)capacity * 4, 256) + 256 + 256;
}

// A fresh blob must be initialised once (clears the header) before the first anyloc_index_add.
extern "C" int anyloc_index_init(void* index, size_t index_bytes, int64_t capacity, int Dv, int normalize, void* stream) {
  ANYLOC_REQUIRE(index, "index_init: null pointer");
  IndexView v;
  if (!carve_index(index, index_bytes, capacity, Dv, normalize, &v)) { set_error("index_init: blob too small"); return ANYLOC_ERR_WORKSPACE; }
  ANYLOC_CHECK_CUDA(cudaMemsetAsync(v.hdr, 0, 256, (cudaStream_t)stream));
  return ANYLOC_OK;
}

extern "C" int anyloc_index_add(void* index, size_t index_bytes, int64_t capacity, int64_t row_offset, const float* rows,
                                int n_rows, int Dv, int normalize, void* stream) {
  ANYLOC_REQUIRE(index && (rows || n_rows == 0), "index_add: null pointer");
  ANYLOC_REQUIRE(n_rows >= 0 && Dv > 0 && Dv % 4 == 0 && row_offset >= 0 && row_offset + n_rows <= capacity,
                 "index_add: bad dims n_rows=%d Dv=%d offset=%lld capacity=%lld", n_rows, Dv, (long long)row_offset,
                 (long long)capacity);
  IndexView v;
  if (!carve_index(index, index_bytes, capacity, Dv, normalize, &v)) {
    set_error("index_add: blob too small (%zu given, %zu needed)", index_bytes, anyloc_index_bytes(capacity, Dv, normalize));
    return ANYLOC_ERR_WORKSPACE;
  }
  if (n_rows == 0) return ANYLOC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(PC_TOPK, st, (v.f16 ? 8.0 : 12.0) * (double)n_rows * Dv);
  const size_t off = (size_t)row_offset * Dv;
  if (v.f16)
    ANYLOC_CHECK_CUDA(launch_normalize_rows<true>(rows, n_rows, Dv, normalize, (__half*)v.hi + off, (__half*)v.lo + off,
                                                  v.sq + row_offset, v.dn + row_offset, v.hdr, st));
  else
    ANYLOC_CHECK_CUDA(launch_normalize_rows<false>(rows, n_rows, Dv, normalize, (float*)v.hi + off, (float*)v.lo + off,
                                                   v.sq + row_offset, nullptr, nullptr, st));
  count_launch();
  return ANYLOC_OK;
}

// Growth: the first n_rows rows of every section (and the header) of `src` -> `dst` (a larger blob of the same Dv / normalize).
extern "C" int anyloc_index_copy(void* dst, size_t dst_bytes, int64_t dst_capacity, const void* src, size_t src_bytes,
                                 int64_t src_capacity, int64_t n_rows, int Dv, int normalize, void* stream) {
  ANYLOC_REQUIRE(dst && src && n_rows >= 0 && n_rows <= src_capacity && n_rows <= dst_capacity, "index_copy: bad arguments");
  IndexView d, s;
  if (!carve_index(dst, dst_bytes, dst_capacity, Dv, normalize, &d) ||
      !carve_index(const_cast<void*>(src), src_bytes, src_capacity, Dv, normalize, &s)) {
    set_error("index_copy: blob too small");
    return ANYLOC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pb = (size_t)n_rows * d.pair_bytes_per_row;
  ANYLOC_CHECK_CUDA(cudaMemcpyAsync(d.hi, s.hi, pb, cudaMemcpyDeviceToDevice, st));
  ANYLOC_CHECK_CUDA(cudaMemcpyAsync(d.lo, s.lo, pb, cudaMemcpyDeviceToDevice, st));
  ANYLOC_CHECK_CUDA(cudaMemcpyAsync(d.sq, s.sq, (size_t)n_rows * 4, cudaMemcpyDeviceToDevice, st));
  ANYLOC_CHECK_CUDA(cudaMemcpyAsync(d.dn, s.dn, (size_t)n_rows * 4, cudaMemcpyDeviceToDevice, st));
  ANYLOC_CHECK_CUDA(cudaMemcpyAsync(d.hdr, s.hdr, 256, cudaMemcpyDeviceToDevice, st));
  return ANYLOC_OK;
}

// workspace of one search: query pairs + |q|^2 + dn_q + the [n_q, n_db] score matrix + candidate lists + flags
extern "C" size_t anyloc_index_search_workspace_bytes(int64_t n_db, int n_q, int Dv, int normalize) {
  const size_t esz = index_uses_f16(Dv, normalize) ? 2 : 4;
  return 2 * align_up((size_t)n_q * Dv * esz, 256) + 2 * align_up((size_t)n_q * 4, 256) +
         align_up((size_t)n_q * (size_t)n_db * 4, 256) + align_up((size_t)n_q * CAND_MAX * 4, 256) +
         align_up((size_t)n_q * 4, 256) + 256 + 1024;
}

extern "C" int anyloc_index_search(const void* index, size_t index_bytes, int64_t capacity, int64_t n_db, const float* qu,
                                   int n_q, int Dv, int k, int metric, int normalize, float* dist, int64_t* idx, void* ws,
                                   size_t ws_bytes, void* stream) {
  ANYLOC_REQUIRE(index && qu && dist && idx && ws, "index_search: null pointer");
  ANYLOC_REQUIRE(n_db > 0 && n_db <= capacity && n_db < (1ll << 31) && n_q >= 0 && Dv > 0 && k > 0,
                 "index_search: bad dims n_db=%lld n_q=%d Dv=%d k=%d", (long long)n_db, n_q, Dv, k);
  ANYLOC_REQUIRE(Dv % 4 == 0, "index_search: Dv=%d must be a multiple of 4", Dv);
  ANYLOC_REQUIRE(metric == ANYLOC_METRIC_IP || metric == ANYLOC_METRIC_L2, "index_search: unknown metric %d", metric);
  if (n_q == 0) return ANYLOC_OK;
  IndexView v;
  if (!carve_index(const_cast<void*>(index), index_bytes, capacity, Dv, normalize, &v)) {
    set_error("index_search: index blob too small");
    return ANYLOC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = v.f16 ? 2 : 4;
  Workspace w(ws, ws_bytes);
  void* qu_hi = w.take<char>((size_t)n_q * Dv * esz);
  void* qu_lo = w.take<char>((size_t)n_q * Dv * esz);
  float* qq = w.take<float>(n_q);
  float* dnq = w.take<float>(n_q);
  float* scores = w.take<float>((size_t)n_q * (size_t)n_db);
  int32_t* cand = w.take<int32_t>((size_t)n_q * CAND_MAX);
  int32_t* cand_n = w.take<int32_t>(n_q);
  int* flags = w.take<int>(64);
  if (!qu_hi || !qu_lo || !qq || !dnq || !scores || !cand || !cand_n || !flags) {
    set_error("index_search: workspace too small (%zu given, %zu needed)", ws_bytes,
              anyloc_index_search_workspace_bytes(n_db, n_q, Dv, normalize));
    return ANYLOC_ERR_WORKSPACE;
  }
  int rc;
  {
    ProfScope ps(PC_TOPK, st, (v.f16 ? 8.0 : 12.0) * (double)n_q * Dv);
    if (v.f16) ANYLOC_CHECK_CUDA(launch_normalize_rows<true>(qu, n_q, Dv, normalize, qu_hi, qu_lo, qq, dnq, nullptr, st));
    else ANYLOC_CHECK_CUDA(launch_normalize_rows<false>(qu, n_q, Dv, normalize, qu_hi, qu_lo, qq, nullptr, nullptr, st));
    count_launch();
  }
  const float alpha = v.f16 ? 1.0f / (kRetrievalScale * kRetrievalScale) : 1.0f;
  const int pair = v.f16 ? ANYLOC_PAIR_F16 : ANYLOC_PAIR_TF32;
  // ---- coarse pass + exact re-scoring of the candidates (inner product, fp16 index); ANYLOC_TOPK_COARSE=0 disables
  static int coarse_env = -1;
  if (coarse_env < 0) { const char* e = getenv("ANYLOC_TOPK_COARSE"); coarse_env = e ? atoi(e) : 1; }
  const bool coarse = coarse_env && v.f16 && metric == ANYLOC_METRIC_IP && k <= 64 && n_db >= 1024 && n_q >= 32;
  const int* gate = nullptr;
  if (coarse) {
    ANYLOC_CHECK_CUDA(cudaMemsetAsync(flags, 0, 256, st));
    // hi-only GEMM: a_lo = b_lo = nullptr selects the single-pass kernels of the tcgen05 engine (recorded under PC_GEMM_TC
    // with the FULL 2*n_q*n_db*Dv algorithmic FLOPs: the coarse pass + re-scoring replace the whole product)
    rc = anyloc_gemm_nt(qu_hi, nullptr, Dv, v.hi, nullptr, Dv, n_q, (int)n_db, Dv, pair, alpha, ANYLOC_EPI_BIAS, nullptr,
                        nullptr, nullptr, scores, nullptr, (int)n_db, ANYLOC_PAIR_TF32, ANYLOC_GEMM_TC3, stream);
    if (rc) return rc;
    ProfScope ps(PC_TOPK, st, 8.0 * (double)n_q * (double)n_db);
    topk_candidates_kernel<<<n_q, 1024, 0, st>>>(scores, (int)n_db, n_db, k, dnq, v.hdr, cand, cand_n, flags);
    ANYLOC_CHECK_LAUNCH();
    topk_rescore_kernel<<<n_q, 256, 0, st>>>((const __half*)v.hi, (const __half*)v.lo, (const __half*)qu_hi,
                                             (const __half*)qu_lo, Dv, k, cand, cand_n, flags, dist, idx);
    ANYLOC_CHECK_LAUNCH();
    gate = flags;            // the exact path below only runs (on the device) if a candidate list overflowed
  }
  // ---- exact path: 3-term score GEMM on the tcgen05 engine (gemm_dispatch records it under PC_GEMM_TC) + selection
  rc = anyloc_gemm_nt_gated(qu_hi, qu_lo, Dv, v.hi, v.lo, Dv, n_q, (int)n_db, Dv, pair, alpha, scores, (int)n_db, gate, stream);
  if (rc) return rc;
  {
    ProfScope ps(PC_TOPK, st, gate ? 0.0 : 8.0 * (double)n_q * (double)n_db);
    topk_select2_kernel<<<n_q, 1024, 0, st>>>(scores, (int)n_db, n_db, k, metric, qq, v.sq, dist, idx, gate);
    ANYLOC_CHECK_LAUNCH();
  }
  return ANYLOC_OK;
}

// One-shot form (get_top_k_recall builds the index and searches it once, utilities.py:449-450): a temporary index in
// the caller's workspace, then the search above.
extern "C" size_t anyloc_topk_workspace_bytes(int n_db, int n_q, int Dv, int k) {
  (void)k;
  // sized for the larger (tf32-pair) layout so that the same buffer serves normalize = 0 / 1
  return anyloc_index_bytes(n_db, (int)align_up((size_t)Dv, 4), 0) +
         anyloc_index_search_workspace_bytes(n_db, n_q, (int)align_up((size_t)Dv, 4), 0) + 512;
}

extern "C" int anyloc_topk(const float* db, const float* qu, int n_db, int n_q, int Dv, int k, int metric,
                           int normalize, float* dist, int64_t* idx, void* ws, size_t ws_bytes,
                           void* stream) {
  ANYLOC_REQUIRE(db && qu && dist && idx && ws, "topk: null pointer");
  ANYLOC_REQUIRE(n_db > 0 && n_q >= 0 && Dv > 0 && k > 0, "topk: bad dims n_db=%d n_q=%d Dv=%d k=%d", n_db, n_q, Dv, k);
  ANYLOC_REQUIRE(Dv % 4 == 0, "topk: Dv=%d must be a multiple of 4", Dv);
  ANYLOC_REQUIRE(metric == ANYLOC_METRIC_IP || metric == ANYLOC_METRIC_L2, "topk: unknown metric %d", metric);
  if (n_q == 0) return ANYLOC_OK;
  const size_t ib = anyloc_index_bytes(n_db, Dv, normalize);
  const size_t need = align_up(ib, 256) + anyloc_index_search_workspace_bytes(n_db, n_q, Dv, normalize);
  if (ws_bytes < need) {
    set_error("topk: workspace too small (%zu given, %zu needed)", ws_bytes, need);
    return ANYLOC_ERR_WORKSPACE;
  }
  int rc = anyloc_index_init(ws, ib, n_db, Dv, normalize, stream);
  if (rc) return rc;
  if ((rc = anyloc_index_add(ws, ib, n_db, 0, db, n_db, Dv, normalize, stream))) return rc;
  char* rest = (char*)ws + align_up(ib, 256);
  return anyloc_index_search(ws, ib, n_db, n_db, qu, n_q, Dv, k, metric, normalize, dist, idx, rest,
                             ws_bytes - align_up(ib, 256), stream);
}
