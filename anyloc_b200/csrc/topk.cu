// Query->database retrieval (reference: /root/reference/utilities.py:435-450, faiss IndexFlatIP /
// IndexFlatL2 exact search).  normalise rows -> score GEMM (fp32-equivalent) -> k-best per query,
// best first, lowest database index first among equal scores.
#include <stdlib.h>
#include "common.cuh"

namespace anyloc {

// fp16-pair scale of unit-norm rows: |s y| <= 4096 < 65504, and s*y - hi stays far above the fp16 subnormal step for
// every element that matters (an element below 2^-12 of the row norm contributes < 2^-36 to a unit dot product).
constexpr float kRetrievalScale = 4096.0f;

// y = x / max(|x|, 1e-12) written as a tf32 (hi,lo) pair (hi+lo == fp32 value); also |y|^2 per row
// (needed by the L2 metric).  One CTA per row.
template <bool F16>
__global__ void __launch_bounds__(256)
normalize_rows_split_kernel(const float* __restrict__ x, int D, int do_norm, void* __restrict__ hi_v,
                            void* __restrict__ lo_v, float* __restrict__ sq) {
  const size_t row = blockIdx.x;
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
  float ss2 = 0.f;
  for (int d = threadIdx.x; d < D4; d += blockDim.x) {
    float4 v = __ldg(xr + d);
    if (do_norm) { v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm; }
    ss2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if constexpr (F16) {
      uint2 h, l;
      split_f16x2(v.x * kRetrievalScale, v.y * kRetrievalScale, h.x, l.x);
      split_f16x2(v.z * kRetrievalScale, v.w * kRetrievalScale, h.y, l.y);
      h2[d] = h; l2[d] = l;
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
}

// k-best selection per query row by k rounds of block arg-best (n_db up to ~1e6, k <= 128).
// metric IP: larger is better; metric L2: score := qq - 2 s + dd, smaller is better.
__global__ void __launch_bounds__(1024)
topk_select_kernel(float* __restrict__ scores, int n_db, int64_t ld, int k, int metric,
                   const float* __restrict__ qq, const float* __restrict__ dd,
                   float* __restrict__ dist, int64_t* __restrict__ idx) {
  const int q = blockIdx.x;
  float* s = scores + (size_t)q * ld;
  if (metric == ANYLOC_METRIC_L2) {
    const float a = qq[q];
    for (int j = threadIdx.x; j < n_db; j += blockDim.x) s[j] = -((a - 2.0f * s[j]) + dd[j]);  // negate: larger=better
    __syncthreads();
  }
  __shared__ float bv[32];
  __shared__ int bi[32];
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY; int besti = 0x7fffffff;
    for (int j = threadIdx.x; j < n_db; j += blockDim.x) {
      float v = s[j];
      if (v > best || (v == best && j < besti)) { best = v; besti = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x < 32) {
      int nw = blockDim.x >> 5;
      best = threadIdx.x < nw ? bv[threadIdx.x] : -INFINITY;
      besti = threadIdx.x < nw ? bi[threadIdx.x] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (threadIdx.x == 0) {
        if (r < n_db && besti != 0x7fffffff) {
          dist[(size_t)q * k + r] = metric == ANYLOC_METRIC_L2 ? -best : best;
          idx[(size_t)q * k + r] = besti;
          s[besti] = -INFINITY;      // exclude from later rounds
        } else {
          dist[(size_t)q * k + r] = metric == ANYLOC_METRIC_L2 ? INFINITY : -INFINITY;
          idx[(size_t)q * k + r] = -1;   // faiss pads with -1 when k > ntotal
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace anyloc

using namespace anyloc;


extern "C" size_t anyloc_topk_workspace_bytes(int n_db, int n_q, int Dv, int k) {
  (void)k;
  size_t Dp = align_up((size_t)Dv, 4);
  return 2 * align_up((size_t)n_db * Dp * 4, 256) + 2 * align_up((size_t)n_q * Dp * 4, 256) +
         align_up((size_t)n_db * 4, 256) + align_up((size_t)n_q * 4, 256) +
         align_up((size_t)n_q * (size_t)n_db * 4, 256) + 4096;
}

extern "C" int anyloc_topk(const float* db, const float* qu, int n_db, int n_q, int Dv, int k, int metric,
                           int normalize, float* dist, int64_t* idx, void* ws, size_t ws_bytes,
                           void* stream) {
  ANYLOC_REQUIRE(db && qu && dist && idx && ws, "topk: null pointer");
  ANYLOC_REQUIRE(n_db > 0 && n_q >= 0 && Dv > 0 && k > 0, "topk: bad dims n_db=%d n_q=%d Dv=%d k=%d", n_db, n_q, Dv, k);
  ANYLOC_REQUIRE(Dv % 4 == 0, "topk: Dv=%d must be a multiple of 4", Dv);
  ANYLOC_REQUIRE(metric == ANYLOC_METRIC_IP || metric == ANYLOC_METRIC_L2, "topk: unknown metric %d", metric);
  if (n_q == 0) return ANYLOC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  float* db_hi = w.take<float>((size_t)n_db * Dv);
  float* db_lo = w.take<float>((size_t)n_db * Dv);
  float* qu_hi = w.take<float>((size_t)n_q * Dv);
  float* qu_lo = w.take<float>((size_t)n_q * Dv);
  float* dd = w.take<float>(n_db);
  float* qq = w.take<float>(n_q);
  float* scores = w.take<float>((size_t)n_q * n_db);
  if (!db_hi || !db_lo || !qu_hi || !qu_lo || !dd || !qq || !scores) {
    set_error("topk: workspace too small (%zu given, %zu needed)", ws_bytes,
              anyloc_topk_workspace_bytes(n_db, n_q, Dv, k));
    return ANYLOC_ERR_WORKSPACE;
  }
  ProfScope ps(PC_TOPK, st, 4.0 * ((double)n_db + n_q) * Dv);
  // Unit-norm rows (the reference's default, norm_descs=True) go through the 2x faster kind::f16 tensor path as fp16
  // pairs of 4096*y; un-normalised rows have no a-priori range and keep the tf32 pairs.  ANYLOC_TOPK_F16=0 disables.
  static int f16_env = -1;
  if (f16_env < 0) { const char* e = getenv("ANYLOC_TOPK_F16"); f16_env = e ? atoi(e) : 1; }
  const bool f16 = f16_env && normalize && (Dv % 8) == 0;
  int rc;
  if (f16) {
    normalize_rows_split_kernel<true><<<n_db, 256, 0, st>>>(db, Dv, normalize, db_hi, db_lo, dd);
    ANYLOC_CHECK_LAUNCH();
    normalize_rows_split_kernel<true><<<n_q, 256, 0, st>>>(qu, Dv, normalize, qu_hi, qu_lo, qq);
    ANYLOC_CHECK_LAUNCH();
    rc = anyloc_gemm_nt(qu_hi, qu_lo, Dv, db_hi, db_lo, Dv, n_q, n_db, Dv, ANYLOC_PAIR_F16,
                        1.0f / (kRetrievalScale * kRetrievalScale), ANYLOC_EPI_BIAS, nullptr, nullptr, nullptr, scores,
                        nullptr, n_db, ANYLOC_PAIR_TF32, ANYLOC_GEMM_AUTO, stream);
  } else {
    normalize_rows_split_kernel<false><<<n_db, 256, 0, st>>>(db, Dv, normalize, db_hi, db_lo, dd);
    ANYLOC_CHECK_LAUNCH();
    normalize_rows_split_kernel<false><<<n_q, 256, 0, st>>>(qu, Dv, normalize, qu_hi, qu_lo, qq);
    ANYLOC_CHECK_LAUNCH();
    rc = anyloc_gemm_nt(qu_hi, qu_lo, Dv, db_hi, db_lo, Dv, n_q, n_db, Dv, ANYLOC_PAIR_TF32, 1.0f, ANYLOC_EPI_BIAS,
                        nullptr, nullptr, nullptr, scores, nullptr, n_db, ANYLOC_PAIR_TF32, ANYLOC_GEMM_AUTO, stream);
  }
  if (rc) return rc;
  topk_select_kernel<<<n_q, 1024, 0, st>>>(scores, n_db, n_db, k, metric, qq, dd, dist, idx);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}
