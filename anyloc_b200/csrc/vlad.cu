// Hard-assignment VLAD (reference: /root/reference/utilities.py:819-926, residuals :956-962,
// assignment fpk.KMeans.predict :849).  See include/anyloc_b200.h for the contract.
//
// v2 pipeline (5 launches per batch):
//   centre_prep : c^_k = c_k/(|c_k|+1e-8) (cosine) or c_k with bias -|c_k|^2/2 (euclid), plus a tf32-rounded copy
//   coarse      : S~[R,K] = X . c^T on the tcgen05 GEMM engine, single tf32 pass straight from the raw fp32
//                 features (no conversion pass; the tensor core truncates) -- HBM-bound, reads X once
//   rescore     : warp per row: |x|, candidate set {k : S~_k >= max - 2 eps} with the rigorous tf32 bound
//                 eps = 2^-9 |x| max|c^|, exact fp32 dot products only for the candidates (row held in registers),
//                 first-max argmax -> labels identical to an exact fp32 evaluation; 1/max(|x|,1e-12)
//   accumulate  : CTA per (image, 128-column slice), warps split the rows, float4 lanes:
//                 sum_{label=k}(x^ - c_k) in shared memory, deterministic per-slice sums of squares
//   normalise   : intra + global L2 normalisation, in place on the [B,K*D] output
// (The v1 FFMA assignment kernel is kept for K > 256 / D > 2048 and as the k-means assignment step.)
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "epilogue.cuh"

namespace anyloc {

// ------------------------------------------------------------------ centre prep
__global__ void vlad_centre_prep_kernel(const float* __restrict__ c, int K, int D, int dist_mode,
                                        float* __restrict__ chat, float* __restrict__ cbias,
                                        float* __restrict__ chat_tf32, float* __restrict__ cnorm,
                                        float* __restrict__ cdnorm /* |c^ - tf32(c^)|, nullable */,
                                        int32_t* __restrict__ zero_a, int zero_a_n, int32_t* __restrict__ zero_b,
                                        int zero_b_n) {
  // the counters of the later launches on this stream are cleared here (saves two memset nodes)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_a_n; i += gridDim.x * blockDim.x) zero_a[i] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_b_n; i += gridDim.x * blockDim.x) zero_b[i] = 0;
  int k = blockIdx.x;
  const float* row = c + (size_t)k * D;
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) { float v = row[d]; ss += v * v; }
  __shared__ float red[32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  ss = red[0];
  float dd = 0.f;                                  // sum (c^ - tf32(c^))^2 over this thread's elements
  if (dist_mode == ANYLOC_DIST_COSINE) {
    const float den = sqrtf(ss) + 1e-8f;            // fpk cos_sim: b / (|b| + 1e-8)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float v = row[d] / den, h, l;
      chat[(size_t)k * D + d] = v;
      if (chat_tf32) { split_tf32(v, h, l); chat_tf32[(size_t)k * D + d] = h; dd += l * l; }
    }
    if (threadIdx.x == 0) { cbias[k] = 0.f; if (cnorm) cnorm[k] = sqrtf(ss) / den; }
  } else {
    // argmax_k 2 x.c_k - |x|^2 - |c_k|^2  ==  argmax_k (x.c_k - |c_k|^2/2)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float v = row[d], h, l;
      chat[(size_t)k * D + d] = v;
      if (chat_tf32) { split_tf32(v, h, l); chat_tf32[(size_t)k * D + d] = h; dd += l * l; }
    }
    if (threadIdx.x == 0) { cbias[k] = -0.5f * ss; if (cnorm) cnorm[k] = sqrtf(ss); }
  }
  if (cdnorm) {
    __syncthreads();
    dd = warp_sum(dd);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dd;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
      cdnorm[k] = sqrtf(tot);
    }
  }
}

// ------------------------------------------------------------------ assign
// One warp handles ROWS rows at a time; lanes stride the feature dimension in float4.
template <int ROWS>
__global__ void __launch_bounds__(256)
vlad_assign_kernel(const float* __restrict__ x, const int32_t* __restrict__ n_valid, int N_per_img,
                   int64_t R, int D, int K, const float* __restrict__ chat,
                   const float* __restrict__ cbias, int32_t* __restrict__ labels,
                   float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int D4 = D >> 2;
  for (int64_t r0 = warp * ROWS; r0 < R; r0 += nwarps * ROWS) {
    const float4* xr[ROWS];
    bool valid[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      int64_t r = r0 + i;
      valid[i] = r < R;
      if (valid[i] && n_valid) {
        int b = (int)(r / N_per_img), n = (int)(r % N_per_img);
        valid[i] = n < n_valid[b];
      }
      xr[i] = reinterpret_cast<const float4*>(x + (valid[i] ? r : r0) * (int64_t)D);
      if (r0 + i >= R) xr[i] = reinterpret_cast<const float4*>(x + r0 * (int64_t)D);
    }
    float ss[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) ss[i] = 0.f;
    for (int d = lane; d < D4; d += 32) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        float4 v = __ldg(xr[i] + d);
        ss[i] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
    float best[ROWS]; int bestk[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) { ss[i] = warp_sum(ss[i]); best[i] = -INFINITY; bestk[i] = 0; }
    for (int k = 0; k < K; ++k) {
      const float4* cr = reinterpret_cast<const float4*>(chat + (size_t)k * D);
      float acc[ROWS];
#pragma unroll
      for (int i = 0; i < ROWS; ++i) acc[i] = 0.f;
      for (int d = lane; d < D4; d += 32) {
        float4 c = __ldg(cr + d);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          float4 v = __ldg(xr[i] + d);     // L1-resident after the norm pass
          acc[i] = fmaf(v.x, c.x, acc[i]); acc[i] = fmaf(v.y, c.y, acc[i]);
          acc[i] = fmaf(v.z, c.z, acc[i]); acc[i] = fmaf(v.w, c.w, acc[i]);
        }
      }
      float bk = cbias[k];
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        float s = warp_sum(acc[i]) + bk;
        if (s > best[i]) { best[i] = s; bestk[i] = k; }   // strict >: lowest index wins ties
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        int64_t r = r0 + i;
        if (r < R) {
          labels[r] = valid[i] ? bestk[i] : -1;
          if (inv_norm) inv_norm[r] = 1.0f / fmaxf(sqrtf(ss[i]), 1e-12f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ rescore (exact labels from coarse scores)
// One warp per row.  The row (D <= 2048) lives in registers; only the candidates whose coarse tf32 score is within
// 2*eps of the row maximum are re-evaluated exactly (fp32 FMA, same c^ as the exact path), so the label equals the
// exact-fp32 argmax (lowest index among exact ties).
template <int MAXV>      // float4 per lane: D <= 128 * MAXV
__global__ void __launch_bounds__(256)
vlad_rescore_kernel(const float* __restrict__ x, const int32_t* __restrict__ n_valid, int N_per_img, int64_t R,
                    int D, int K, const float* __restrict__ chat, const float* __restrict__ cbias,
                    const float* __restrict__ cnorm, const float* __restrict__ coarse /*[R,K]*/,
                    int32_t* __restrict__ labels, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= R) return;
  bool valid = true;
  if (n_valid) { int b = (int)(row / N_per_img), n = (int)(row % N_per_img); valid = n < n_valid[b]; }
  const int D4 = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * (int64_t)D);
  float4 v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int d = lane + i * 32;
    if (d < D4) { v[i] = __ldg(xr + d); ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w; }
  }
  ss = warp_sum(ss);
  const float xn = sqrtf(ss);
  // coarse maximum and the largest centre norm (lanes stride over k)
  float smax = -INFINITY, cmax = 0.f;
  for (int k = lane; k < K; k += 32) {
    smax = fmaxf(smax, coarse[row * K + k]);
    cmax = fmaxf(cmax, cnorm[k]);
  }
  smax = warp_max(smax); cmax = warp_max(cmax);
  // |S~_k - S_k| <= (2^-10 + 2^-11) sum|x_i c_i| <= 1.5 * 2^-10 |x||c_k| < 2^-9 |x||c_k|  (truncated x, rounded c)
  const float thresh = smax - 2.0f * (0.001953125f * xn * cmax) - 1e-30f;
  float best = -INFINITY; int bestk = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    const bool cand = k < K && coarse[row * K + k] >= thresh;
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    while (mask) {
      // up to four candidates at a time (independent accumulators hide the L1/L2 latency of the centre rows)
      int kk[4]; int n = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (mask) { kk[q] = k0 + __ffs(mask) - 1; mask &= mask - 1; ++n; } else kk[q] = kk[0];
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        int d = lane + i * 32;
        if (d < D4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 c = __ldg(reinterpret_cast<const float4*>(chat + (size_t)kk[q] * D) + d);
            acc[q] = fmaf(v[i].x, c.x, acc[q]); acc[q] = fmaf(v[i].y, c.y, acc[q]);
            acc[q] = fmaf(v[i].z, c.z, acc[q]); acc[q] = fmaf(v[i].w, c.w, acc[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sc = warp_sum(acc[q]) + cbias[kk[q]];
        if (q < n && sc > best) { best = sc; bestk = kk[q]; }   // ascending k, strict >: lowest index wins exact ties
      }
    }
  }
  if (lane == 0) {
    labels[row] = valid ? bestk : -1;
    if (inv_norm) inv_norm[row] = 1.0f / fmaxf(xn, 1e-12f);
  }
}

// ------------------------------------------------------------------ accumulate v2
// CTA = (image, 128-column slice); WARPS warps split the rows; lane owns 4 consecutive columns (float4).
// Per-warp accumulators [K][128] in shared memory (no conflicts: a warp touches 512 contiguous bytes per row),
// reduced across warps at the end in a fixed order -> deterministic.
__global__ void __launch_bounds__(256, 2)
vlad_accumulate2_kernel(const float* __restrict__ x, const int32_t* __restrict__ labels,
                        const float* __restrict__ inv_norm, const float* __restrict__ centers,
                        int N, int D, int K, int norm_descs, int warps, float* __restrict__ vlad,
                        float* __restrict__ partial_ss /* [B,K,nslices] */) {
  extern __shared__ float sm[];
  float* cen = sm;                                        // [K][128]
  float* acc = sm + (size_t)K * 128;                      // [warps][K][128]
  const int b = blockIdx.y, slice = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int col = slice * 128 + lane * 4;
  const bool colok = col < D;                             // D % 4 == 0
  for (int i = t; i < K * 32; i += blockDim.x) {
    const int k = i >> 5, c4 = (i & 31) * 4, gc = slice * 128 + c4;
    float4 cv = gc < D ? __ldg(reinterpret_cast<const float4*>(centers + (size_t)k * D + gc)) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(cen + k * 128 + c4) = cv;
  }
  for (int i = t; i < warps * K * 32; i += blockDim.x) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (w < warps && colok) {
    float* my = acc + (size_t)w * K * 128 + lane * 4;
    const float* xb = x + (size_t)b * N * D + col;
    const int32_t* lb = labels + (size_t)b * N;
    const float* ib = inv_norm + (size_t)b * N;
    constexpr int U = 16;                                 // rows in flight per warp (latency hiding)
    for (int n0 = w; n0 < N; n0 += U * warps) {
      float4 v[U]; int lab[U]; float sc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int n = n0 + u * warps;
        lab[u] = n < N ? lb[n] : -1;
        sc[u] = (n < N && norm_descs) ? ib[n] : 1.0f;
        v[u] = lab[u] >= 0 ? __ldg(reinterpret_cast<const float4*>(xb + (size_t)n * D)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (lab[u] >= 0) {
          float4 a = *reinterpret_cast<float4*>(my + lab[u] * 128);
          const float4 c = *reinterpret_cast<const float4*>(cen + lab[u] * 128 + lane * 4);
          a.x += v[u].x * sc[u] - c.x; a.y += v[u].y * sc[u] - c.y; a.z += v[u].z * sc[u] - c.z; a.w += v[u].w * sc[u] - c.w;
          *reinterpret_cast<float4*>(my + lab[u] * 128) = a;
        }
      }
    }
  }
  __syncthreads();
  // reduce the warps' partials (fixed order), write V and the per-slice sums of squares
  const int nslices = gridDim.x;
  for (int k = w; k < K; k += blockDim.x >> 5) {          // one warp per cluster row
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < warps; ++q) {
      float4 p = *reinterpret_cast<const float4*>(acc + ((size_t)q * K + k) * 128 + lane * 4);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    float ss = 0.f;
    if (colok) {
      *reinterpret_cast<float4*>(vlad + ((size_t)b * K + k) * D + col) = a;
      ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = warp_sum(ss);
    if (lane == 0) partial_ss[((size_t)b * K + k) * nslices + slice] = ss;
  }
}

// ------------------------------------------------------------------ accumulate v3 (+ fused normalisation)
// CTA = (128-column slice, image), 8 warps, 4-5 CTAs per SM (one wave at the BASELINE shapes).  The image's rows are
// counting-sorted by label in shared memory (stable: rows of a cluster stay in row order), the sorted list is cut into
// tasks of <= 64 rows of ONE cluster, and warps grab tasks dynamically: 8 independent 512-byte row segments in flight
// per warp, the partial sum lives in registers (lane = 4 columns) -- no shared-memory accumulators, no
// read-modify-write chains, and a skewed vocabulary (one cluster holding a third of the image) no longer serialises on
// one warp.  Clusters of several tasks are combined from shared-memory slots in task order, so every sum has ONE fixed
// order: bitwise reproducible, batch == single image.  The last CTA of an image to finish (global ticket) applies the
// intra- and global L2 normalisation to that image's descriptor while it is still in L2 (no separate launch).
constexpr int ACC3_WARPS = 8;
constexpr int ACC3_SEG = 64;
static inline int acc3_max_tasks(int N, int K) { return N / ACC3_SEG + K + 1; }
__device__ __forceinline__ int acc3_max_tasks_dev(int N, int K) { return N / ACC3_SEG + K + 1; }
static inline int acc3_max_slots(int N) { return 2 * (N / ACC3_SEG) + 2; }
static inline size_t acc3_smem_bytes(int N, int K) {
  return ((size_t)4 * N + 3 * (size_t)(K + 1) + (size_t)ACC3_WARPS * K + 2 * (size_t)K + acc3_max_tasks(N, K) + 4) * 4 +
         (size_t)acc3_max_slots(N) * 512;
}
__global__ void __launch_bounds__(ACC3_WARPS * 32, 4)
vlad_accumulate3_kernel(const float* __restrict__ x, const int32_t* __restrict__ labels,
                        const float* __restrict__ inv_norm, const float* __restrict__ centers, int N, int D, int K,
                        int norm_descs, int intra_norm, float* vlad, float* partial_ss /* [B,K,nslices] */,
                        int32_t* done /* [B], zero on entry */, int32_t* reset_ctr /* nullable */, int prefetch, int wait_all,
                        unsigned long long* dbg /* ANYLOC_VLAD_TIMELINE only: 8 ns stamps per CTA, nullable */) {
  auto stamp = [&](int i) {
    if (dbg && threadIdx.x == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + i] = t;
    }
  };
  stamp(0);
  extern __shared__ __align__(16) int sm3[];
  // prepared-vocabulary calls: the work-list counter of the assignment stage (already consumed on this stream) is
  // cleared here for the next call, so no launch is spent on it
  if (reset_ctr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *reset_ctr = 0;
  int* ooff = sm3;                                          // [N] n * D of the rows, sorted by label (stable)
  int* lab = ooff + N;                                      // [N]
  float* inv_s = reinterpret_cast<float*>(lab + N);         // [N] 1/|x| in the same sorted order
  int* start = reinterpret_cast<int*>(inv_s + N);           // [K+1] first sorted position of cluster k
  int* tstart = start + K + 1;                              // [K+1] first task of cluster k
  int* sbase = tstart + K + 1;                              // [K+1] first partial-sum slot of a multi-task cluster
  int* cntw = sbase + K + 1;                                // [ACC3_WARPS][K]
  float* kss = reinterpret_cast<float*>(cntw + ACC3_WARPS * K);   // [K]
  float* ksq = kss + K;                                     // [K]
  int* task_k = reinterpret_cast<int*>(ksq + K);            // [max_tasks]
  float* inv = reinterpret_cast<float*>(task_k + acc3_max_tasks_dev(N, K));   // [N] 1/|x| in row order (prologue only)
  float* slots = reinterpret_cast<float*>(sm3) +
                 (((size_t)4 * N + 3 * (size_t)(K + 1) + (size_t)ACC3_WARPS * K + 2 * (size_t)K + acc3_max_tasks_dev(N, K) + 3) & ~(size_t)3);
  __shared__ int next_task, s_last;
  __shared__ float s_gnorm;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // images in reverse order: the assignment pass streamed them in ascending order, so the last ones are the most
  // likely to still sit in L2 when this kernel starts
  const int b = (int)gridDim.y - 1 - (int)blockIdx.y, slice = blockIdx.x, nslices = gridDim.x;
  const int col = slice * 128 + lane * 4;
  const bool colok = col < D;                               // D % 4 == 0
  for (int n = t; n < N; n += blockDim.x) {
    lab[n] = labels[(size_t)b * N + n];
    inv[n] = norm_descs ? inv_norm[(size_t)b * N + n] : 1.0f;
  }
  for (int i = t; i < ACC3_WARPS * K; i += blockDim.x) cntw[i] = 0;
  if (t == 0) next_task = 0;
  __syncthreads();
  stamp(1);
  // per-warp histograms over contiguous row chunks
  const int chunk = (((N + ACC3_WARPS - 1) / ACC3_WARPS) + 31) & ~31;
  const int r0 = min(N, w * chunk), r1 = min(N, r0 + chunk);
  for (int n = r0 + lane; n < r1; n += 32) { const int l = lab[n]; if (l >= 0) atomicAdd(&cntw[w * K + l], 1); }
  __syncthreads();
  for (int k = t; k < K; k += blockDim.x) {                 // exclusive prefix over the warps, cluster totals
    int tot = 0;
    for (int ww = 0; ww < ACC3_WARPS; ++ww) { const int c = cntw[ww * K + k]; cntw[ww * K + k] = tot; tot += c; }
    start[k] = tot;
  }
  __syncthreads();
  if (w == 0) {                                             // exclusive scans: rows, tasks, partial-sum slots
    int run_r = 0, run_t = 0, run_s = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      const int c = k < K ? start[k] : 0;
      const int nt = k < K ? max(1, (c + ACC3_SEG - 1) / ACC3_SEG) : 0;
      const int ns = nt > 1 ? nt : 0;
      int ir = c, it = nt, is = ns;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int yr = __shfl_up_sync(0xffffffffu, ir, o), yt = __shfl_up_sync(0xffffffffu, it, o),
                  ys = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) { ir += yr; it += yt; is += ys; }
      }
      if (k < K) { start[k] = run_r + ir - c; tstart[k] = run_t + it - nt; sbase[k] = run_s + is - ns; }
      run_r += __shfl_sync(0xffffffffu, ir, 31);
      run_t += __shfl_sync(0xffffffffu, it, 31);
      run_s += __shfl_sync(0xffffffffu, is, 31);
    }
    if (lane == 0) { start[K] = run_r; tstart[K] = run_t; sbase[K] = run_s; }
  }
  __syncthreads();
  for (int k = t; k < K; k += blockDim.x)                   // task table
    for (int q = tstart[k]; q < tstart[k + 1]; ++q) task_k[q] = k;
  for (int n0 = r0; n0 < r1; n0 += 32) {                    // stable placement
    const int n = n0 + lane;
    const int l = n < r1 ? lab[n] : -1;
    const bool active = l >= 0;
    const float iv = active ? inv[n] : 1.0f;
    const unsigned am = __ballot_sync(0xffffffffu, active);
    unsigned peers = 0; int rank = 0;
    if (active) {
      peers = __match_any_sync(am, l);
      rank = __popc(peers & ((1u << lane) - 1u));
      const int pos = start[l] + cntw[w * K + l] + rank;
      ooff[pos] = n * D;
      inv_s[pos] = iv;
    }
    __syncwarp();
    if (active && rank == 0) cntw[w * K + l] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  stamp(2);
  // tasks -> registers.  Every warp grabs its NEXT task one task early; with `prefetch` it also bulk-prefetches that
  // task's row segments into L2 (cp.async.bulk.prefetch.L2: no destination registers).  Measured neutral: this phase
  // already streams at 5.2-5.4 TB/s (profiles/r01_vlad_v3.md, section 4b), so the prefetch is off by default.
  const float* xb = x + (size_t)b * N * D + col;
  const float* xs = x + (size_t)b * N * D + slice * 128;                     // this slice, lane-independent
  const uint32_t rowbytes = (uint32_t)min(128, D - slice * 128) * 4u;
  const int ntasks = tstart[K];
  auto grab = [&]() { int q = 0; if (lane == 0) q = atomicAdd(&next_task, 1); return __shfl_sync(0xffffffffu, q, 0); };
  auto prefetch_task = [&](int q) {
    if (q >= ntasks || !prefetch) return;
    const int k = task_k[q];
    const int s = start[k] + (q - tstart[k]) * ACC3_SEG, e = min(start[k + 1], s + ACC3_SEG);
    for (int i = s + lane; i < e; i += 32)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xs + ooff[i]), "r"(rowbytes) : "memory");
  };
  int q = grab();
  prefetch_task(q);
  while (q < ntasks) {
    const int qn = grab();
    prefetch_task(qn);
    const int k = task_k[q];
    const int seg = q - tstart[k], nt = tstart[k + 1] - tstart[k];
    const int s = start[k] + seg * ACC3_SEG, e = min(start[k + 1], s + ACC3_SEG);
    const float4 c = colok ? __ldg(reinterpret_cast<const float4*>(centers + (size_t)k * D + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colok) {
      // The in-flight bytes live in registers (8 x 16 B per lane = 4 KB per warp; 4 CTAs x 8 warps -> 128 KB per SM),
      // so the loop is kept lean: row offsets and 1/|x| were laid out in sorted order by the placement pass.
      constexpr int U = 8;
      int i = s;
      for (; i + U <= e; i += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(xb + ooff[i + u]));
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float sc = inv_s[i + u];
          a.x += v[u].x * sc - c.x; a.y += v[u].y * sc - c.y; a.z += v[u].z * sc - c.z; a.w += v[u].w * sc - c.w;
        }
      }
      if (i < e) {                                           // tail: < U rows, same order
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u < e) v[u] = __ldg(reinterpret_cast<const float4*>(xb + ooff[i + u]));
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (i + u < e) {
            const float sc = inv_s[i + u];
            a.x += v[u].x * sc - c.x; a.y += v[u].y * sc - c.y; a.z += v[u].z * sc - c.z; a.w += v[u].w * sc - c.w;
          }
        }
      }
    }
    if (nt == 1) {
      if (colok) *reinterpret_cast<float4*>(vlad + ((size_t)b * K + k) * D + col) = a;
      const float ss = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
      if (lane == 0) kss[k] = ss;
    } else {
      *reinterpret_cast<float4*>(slots + (size_t)(sbase[k] + seg) * 128 + lane * 4) = a;
    }
    q = qn;
  }
  __syncthreads();
  stamp(3);
  for (int k = w; k < K; k += ACC3_WARPS) {                 // clusters of several tasks: combine in task order
    const int nt = tstart[k + 1] - tstart[k];
    if (nt <= 1) continue;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < nt; ++q) {
      const float4 p = *reinterpret_cast<const float4*>(slots + (size_t)(sbase[k] + q) * 128 + lane * 4);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    if (colok) *reinterpret_cast<float4*>(vlad + ((size_t)b * K + k) * D + col) = a;
    const float ss = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
    if (lane == 0) kss[k] = ss;
  }
  __syncthreads();
  for (int k = t; k < K; k += blockDim.x) partial_ss[((size_t)b * K + k) * nslices + slice] = kss[k];
  __syncthreads();
  if (t == 0) {
    __threadfence();      // cumulative: orders every write the barrier above made visible to this thread (the pattern of
                          // cooperative-groups grid sync), instead of 256 per-thread fences
    s_last = (atomicAdd(&done[b], 1) == nslices - 1);
  }
  __syncthreads();
  stamp(4);
  if (wait_all) {
    // Whole grid co-resident (checked on the host): every slice-CTA waits until all slices of its image have published
    // their sums of squares, derives the SAME scales in the same order, and normalises ITS OWN 128-column slice -- the
    // normalisation is spread over all CTAs of the image instead of serialising K*D elements behind the last one
    // (measured tail of the last-CTA variant: 10 us of 41 at c2, 30-40 us of 130 at c5).
    if (t == 0) {
      const long long t0 = clock64();
      for (;;) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(done + b) : "memory");
        if (v >= nslices) break;
        __nanosleep(64);
        if (clock64() - t0 > 8000000000LL) __trap();      // never hang the GPU
      }
    }
    __syncthreads();
    for (int k = t; k < K; k += blockDim.x) {
      float ss = 0.f;
      for (int s = 0; s < nslices; ++s) ss += __ldcg(partial_ss + ((size_t)b * K + k) * nslices + s);
      const float nk = sqrtf(ss);
      const float sc = intra_norm ? 1.0f / fmaxf(nk, 1e-12f) : 1.0f;
      kss[k] = sc;
      const float nb = nk * sc;
      ksq[k] = nb * nb;
    }
    __syncthreads();
    if (t == 0) {
      float tot = 0.f;
      for (int k = 0; k < K; ++k) tot += ksq[k];
      s_gnorm = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    }
    __syncthreads();
    const float g = s_gnorm;
    const int nq = K * 32;                                  // float4 elements of this slice
    constexpr int UW = 8;
    for (int i0 = t; i0 < nq; i0 += blockDim.x * UW) {
      float4 v[UW];
#pragma unroll
      for (int u = 0; u < UW; ++u) {
        const int i = i0 + u * blockDim.x, c4 = slice * 128 + (i & 31) * 4;
        if (i < nq && c4 < D) v[u] = __ldcg(reinterpret_cast<const float4*>(vlad + ((size_t)b * K + (i >> 5)) * D + c4));
      }
#pragma unroll
      for (int u = 0; u < UW; ++u) {
        const int i = i0 + u * blockDim.x, c4 = slice * 128 + (i & 31) * 4;
        if (i < nq && c4 < D) {
          const float sc = kss[i >> 5];
          v[u].x = (v[u].x * sc) * g; v[u].y = (v[u].y * sc) * g; v[u].z = (v[u].z * sc) * g; v[u].w = (v[u].w * sc) * g;
          *reinterpret_cast<float4*>(vlad + ((size_t)b * K + (i >> 5)) * D + c4) = v[u];
        }
      }
    }
    __syncthreads();
    stamp(5);
    return;
  }
  if (!s_last) return;
  // ---- last CTA of this image: intra- and global normalisation (same arithmetic as vlad_normalize_kernel)
  __threadfence();
  for (int k = t; k < K; k += blockDim.x) {
    float ss = 0.f;
    for (int s = 0; s < nslices; ++s) ss += __ldcg(partial_ss + ((size_t)b * K + k) * nslices + s);
    const float nk = sqrtf(ss);
    const float sc = intra_norm ? 1.0f / fmaxf(nk, 1e-12f) : 1.0f;
    kss[k] = sc;
    const float nb = nk * sc;
    ksq[k] = nb * nb;
  }
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int k = 0; k < K; ++k) tot += ksq[k];
    s_gnorm = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    done[b] = 0;
  }
  __syncthreads();
  const float g = s_gnorm;
  float4* vb = reinterpret_cast<float4*>(vlad + (size_t)b * K * D);
  const int D4 = D >> 2, total4 = K * D4;
  constexpr int UN = 8;                                     // loads batched ahead of the stores (L2 latency chain)
  for (int i0 = t; i0 < total4; i0 += blockDim.x * UN) {
    float4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total4) v[u] = __ldcg(vb + i);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total4) {
        const float sc = kss[i / D4];
        // two separate multiplications like F.normalize(intra) then F.normalize(global)
        v[u].x = (v[u].x * sc) * g; v[u].y = (v[u].y * sc) * g; v[u].z = (v[u].z * sc) * g; v[u].w = (v[u].w * sc) * g;
        vb[i] = v[u];
      }
    }
  }
  __syncthreads();
  stamp(5);
}

// ------------------------------------------------------------------ accumulate
constexpr int ACC_COLS = 128;
__global__ void __launch_bounds__(ACC_COLS)
vlad_accumulate_kernel(const float* __restrict__ x, const int32_t* __restrict__ labels,
                       const float* __restrict__ inv_norm, const float* __restrict__ centers,
                       int N, int D, int K, int norm_descs, float* __restrict__ vlad,
                       float* __restrict__ partial_ss /* [B,K,nslices] */) {
  extern __shared__ float sm[];
  float* acc = sm;                    // [K][ACC_COLS]
  float* cen = sm + (size_t)K * ACC_COLS;  // [K][ACC_COLS]
  int* lab = reinterpret_cast<int*>(cen + (size_t)K * ACC_COLS);   // [N]
  float* inv = reinterpret_cast<float*>(lab + N);                  // [N]
  const int b = blockIdx.y, slice = blockIdx.x, t = threadIdx.x;
  const int col = slice * ACC_COLS + t;
  const bool colok = col < D;
  for (int k = 0; k < K; ++k) {
    acc[k * ACC_COLS + t] = 0.f;
    cen[k * ACC_COLS + t] = colok ? centers[(size_t)k * D + col] : 0.f;
  }
  for (int n = t; n < N; n += ACC_COLS) {
    lab[n] = labels[(size_t)b * N + n];
    inv[n] = norm_descs ? inv_norm[(size_t)b * N + n] : 1.0f;
  }
  __syncthreads();
  const float* xb = x + (size_t)b * N * D + col;
  if (colok) {
    int n = 0;
    for (; n + 4 <= N; n += 4) {
      float v0 = __ldg(xb + (size_t)(n + 0) * D), v1 = __ldg(xb + (size_t)(n + 1) * D);
      float v2 = __ldg(xb + (size_t)(n + 2) * D), v3 = __ldg(xb + (size_t)(n + 3) * D);
      int l0 = lab[n], l1 = lab[n + 1], l2 = lab[n + 2], l3 = lab[n + 3];
      if (l0 >= 0) acc[l0 * ACC_COLS + t] += v0 * inv[n + 0] - cen[l0 * ACC_COLS + t];
      if (l1 >= 0) acc[l1 * ACC_COLS + t] += v1 * inv[n + 1] - cen[l1 * ACC_COLS + t];
      if (l2 >= 0) acc[l2 * ACC_COLS + t] += v2 * inv[n + 2] - cen[l2 * ACC_COLS + t];
      if (l3 >= 0) acc[l3 * ACC_COLS + t] += v3 * inv[n + 3] - cen[l3 * ACC_COLS + t];
    }
    for (; n < N; ++n) {
      float v = __ldg(xb + (size_t)n * D);
      int l = lab[n];
      if (l >= 0) acc[l * ACC_COLS + t] += v * inv[n] - cen[l * ACC_COLS + t];
    }
  }
  __syncthreads();
  // write un-normalised V and the per-slice sum of squares (warp 0..3 -> fixed order reduce)
  __shared__ float red[ACC_COLS / 32];
  const int nslices = gridDim.x;
  for (int k = 0; k < K; ++k) {
    float v = acc[k * ACC_COLS + t];
    if (colok) vlad[((size_t)b * K + k) * D + col] = v;
    float s = warp_sum(colok ? v * v : 0.f);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    if (t == 0) {
      float tot = 0.f;
      for (int w = 0; w < ACC_COLS / 32; ++w) tot += red[w];
      partial_ss[((size_t)b * K + k) * nslices + slice] = tot;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ soft assignment (utilities.py:862-887)
// a[r,k] = softmax_k(temp * cos(x_r, c_k)) with F.cosine_similarity's clamps (each norm clamped at 1e-8).
// One warp handles ROWS rows per pass over the centres (rows stay L1-resident); scores are staged in shared memory.
template <int ROWS>
__global__ void __launch_bounds__(256)
vlad_soft_assign_kernel(const float* __restrict__ x, const int32_t* __restrict__ n_valid, int N_per_img, int64_t R,
                        int D, int K, const float* __restrict__ chat /* c / max(|c|, 1e-8) */, float temp,
                        float* __restrict__ assign /*[R,K]*/, float* __restrict__ inv_norm) {
  extern __shared__ float sc_smem[];                 // [warps][ROWS][K]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* sc = sc_smem + (size_t)wib * ROWS * K;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int D4 = D >> 2;
  for (int64_t r0 = warp * ROWS; r0 < R; r0 += nwarps * ROWS) {
    const float4* xr[ROWS];
    bool valid[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      int64_t r = r0 + i;
      valid[i] = r < R;
      if (valid[i] && n_valid) valid[i] = (int)(r % N_per_img) < n_valid[(int)(r / N_per_img)];
      xr[i] = reinterpret_cast<const float4*>(x + (r < R ? r : r0) * (int64_t)D);
    }
    float ss[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) ss[i] = 0.f;
    for (int d = lane; d < D4; d += 32) {
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        float4 v = __ldg(xr[i] + d);
        ss[i] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
    float rx[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) { ss[i] = sqrtf(warp_sum(ss[i])); rx[i] = temp / fmaxf(ss[i], 1e-8f); }
    for (int k = 0; k < K; ++k) {
      const float4* cr = reinterpret_cast<const float4*>(chat + (size_t)k * D);
      float acc[ROWS];
#pragma unroll
      for (int i = 0; i < ROWS; ++i) acc[i] = 0.f;
      for (int d = lane; d < D4; d += 32) {
        float4 c = __ldg(cr + d);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          float4 v = __ldg(xr[i] + d);
          acc[i] = fmaf(v.x, c.x, acc[i]); acc[i] = fmaf(v.y, c.y, acc[i]);
          acc[i] = fmaf(v.z, c.z, acc[i]); acc[i] = fmaf(v.w, c.w, acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        float s = warp_sum(acc[i]) * rx[i];
        if (lane == 0) sc[i * K + k] = s;
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      int64_t r = r0 + i;
      if (r >= R) continue;
      float m = -INFINITY;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, sc[i * K + k]);
      m = warp_max(m);
      float z = 0.f;
      for (int k = lane; k < K; k += 32) { float e = expf(sc[i * K + k] - m); sc[i * K + k] = e; z += e; }
      z = warp_sum(z);
      const float iz = valid[i] ? 1.0f / z : 0.f;       // padded rows of ragged batches carry no weight
      for (int k = lane; k < K; k += 32) assign[r * K + k] = sc[i * K + k] * iz;
      if (lane == 0 && inv_norm) inv_norm[r] = 1.0f / fmaxf(ss[i], 1e-12f);
    }
    __syncwarp();
  }
}

// V[b,k,:] = sum_q a[q,k] * sum_c (x^_q - c_c) = K * sum_q a[q,k] x^_q - (sum_q a[q,k]) * sum_c c_c
// (the reference sums cluster k's weight over the residuals to ALL centres, utilities.py:881-884).
// CTA = (128-column slice, image); thread = column; KC cluster accumulators in registers per pass.
constexpr int SOFT_KC = 32, SOFT_QT = 64;
__global__ void __launch_bounds__(ACC_COLS)
vlad_soft_accumulate_kernel(const float* __restrict__ x, const float* __restrict__ assign,
                            const float* __restrict__ inv_norm, const float* __restrict__ centers, int N, int D, int K,
                            int norm_descs, float* __restrict__ vlad, float* __restrict__ partial_ss) {
  __shared__ __align__(16) float a_tile[SOFT_QT][SOFT_KC];
  __shared__ float inv_tile[SOFT_QT];
  __shared__ float red[ACC_COLS / 32][SOFT_KC];
  const int t = threadIdx.x, slice = blockIdx.x, b = blockIdx.y, nslices = gridDim.x;
  const int col = slice * ACC_COLS + t;
  const bool colok = col < D;
  const float* xb = x + (size_t)b * N * D;
  const float* ab = assign + (size_t)b * N * K;
  float csum = 0.f;
  if (colok) for (int c = 0; c < K; ++c) csum += __ldg(centers + (size_t)c * D + col);
  for (int k0 = 0; k0 < K; k0 += SOFT_KC) {
    const int kc = min(SOFT_KC, K - k0);
    float acc[SOFT_KC];
#pragma unroll
    for (int j = 0; j < SOFT_KC; ++j) acc[j] = 0.f;
    float wsum = 0.f;                         // thread j < kc: sum_q a[q, k0 + j]
    for (int q0 = 0; q0 < N; q0 += SOFT_QT) {
      const int qn = min(SOFT_QT, N - q0);
      __syncthreads();
      for (int i = t; i < SOFT_QT * SOFT_KC; i += ACC_COLS) {
        int q = i / SOFT_KC, j = i % SOFT_KC;
        a_tile[q][j] = (q < qn && j < kc) ? __ldg(ab + (size_t)(q0 + q) * K + k0 + j) : 0.f;
      }
      for (int q = t; q < SOFT_QT; q += ACC_COLS)
        inv_tile[q] = (q < qn) ? (norm_descs ? inv_norm[(size_t)b * N + q0 + q] : 1.0f) : 0.f;
      __syncthreads();
      if (t < SOFT_KC) for (int q = 0; q < qn; ++q) wsum += a_tile[q][t];
      if (colok) {
#pragma unroll 4
        for (int q = 0; q < qn; ++q) {
          const float xv = __ldg(xb + (size_t)(q0 + q) * D + col) * inv_tile[q];
          const float4* ar = reinterpret_cast<const float4*>(a_tile[q]);
#pragma unroll
          for (int j4 = 0; j4 < SOFT_KC / 4; ++j4) {
            float4 a = ar[j4];
            acc[4 * j4 + 0] = fmaf(a.x, xv, acc[4 * j4 + 0]); acc[4 * j4 + 1] = fmaf(a.y, xv, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(a.z, xv, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(a.w, xv, acc[4 * j4 + 3]);
          }
        }
      }
    }
    // epilogue for this cluster chunk: values, then per-(image, cluster, slice) sums of squares
    __syncthreads();
    if (t < SOFT_KC) inv_tile[t] = wsum;      // SOFT_KC <= SOFT_QT: reuse as the weight sums
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SOFT_KC; ++j) {
      float v = 0.f;
      if (colok && j < kc) {
        v = (float)K * acc[j] - inv_tile[j] * csum;
        vlad[((size_t)b * K + k0 + j) * D + col] = v;
      }
      float sq = warp_sum(v * v);
      if ((t & 31) == 0) red[t >> 5][j] = sq;
    }
    __syncthreads();
    if (t < kc) {
      float tot = 0.f;
      for (int w = 0; w < ACC_COLS / 32; ++w) tot += red[w][t];
      partial_ss[((size_t)b * K + k0 + t) * nslices + slice] = tot;
    }
  }
}

// ------------------------------------------------------------------ normalise
__global__ void __launch_bounds__(256)
vlad_normalize_kernel(float* __restrict__ vlad, const float* __restrict__ partial_ss, int D, int K,
                      int nslices, int intra_norm) {
  extern __shared__ float scale[];   // [K]
  __shared__ float gnorm;
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float ss = 0.f;
    for (int s = 0; s < nslices; ++s) ss += partial_ss[((size_t)b * K + k) * nslices + s];
    float nk = sqrtf(ss);
    float sc = intra_norm ? 1.0f / fmaxf(nk, 1e-12f) : 1.0f;
    scale[k] = sc;
    // squared norm of the block after intra-normalisation
    float nb = nk * sc;
    scale[K + k] = nb * nb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < K; ++k) tot += scale[K + k];
    gnorm = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
  }
  __syncthreads();
  float* v = vlad + (size_t)b * K * D;
  const size_t total = (size_t)K * D;
  for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.y * blockDim.x) {
    int k = (int)(i / D);
    // two separate multiplications like F.normalize(intra) then F.normalize(global)
    v[i] = (v[i] * scale[k]) * gnorm;
  }
}

// ------------------------------------------------------------------ k-means update
// Deterministic: every (column slice, row chunk) CTA writes ITS partial sums / counts, the finalize kernel adds the
// chunks in chunk order -- no floating-point atomics, so a fitted vocabulary is bit-reproducible for a fixed seed
// (like the reference's single-threaded mask @ X).
__global__ void kmeans_accumulate_kernel(const float* __restrict__ x, const int32_t* __restrict__ labels,
                                         int64_t R, int D, int K, float* __restrict__ psums /* [chunks,K,D] */,
                                         float* __restrict__ pcounts /* [chunks,K] */) {
  // grid (D/128 slices, row-chunks); shared [K][128] partial sums
  extern __shared__ float acc[];
  const int t = threadIdx.x, col = blockIdx.x * ACC_COLS + t;
  const bool colok = col < D;
  for (int k = 0; k < K; ++k) acc[k * ACC_COLS + t] = 0.f;
  float* cnt = acc + (size_t)K * ACC_COLS;
  for (int k = t; k < K; k += ACC_COLS) cnt[k] = 0.f;
  __syncthreads();
  int64_t rows_per = (R + gridDim.y - 1) / gridDim.y;
  int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  for (int64_t r = r0; r < r1; ++r) {
    int l = labels[r];
    if (l < 0) continue;
    if (colok) acc[l * ACC_COLS + t] += __ldg(x + r * D + col);
    if (blockIdx.x == 0 && t == 0) cnt[l] += 1.f;
  }
  __syncthreads();
  float* ps = psums + (size_t)blockIdx.y * K * D;
  for (int k = 0; k < K; ++k)
    if (colok) ps[(size_t)k * D + col] = acc[k * ACC_COLS + t];
  if (blockIdx.x == 0)
    for (int k = t; k < K; k += ACC_COLS) pcounts[(size_t)blockIdx.y * K + k] = cnt[k];
}

__global__ void __launch_bounds__(256)
kmeans_finalize_kernel(const float* __restrict__ psums, const float* __restrict__ pcounts, int chunks,
                       const float* __restrict__ old_c, int D, int K, float* __restrict__ new_c,
                       float* __restrict__ perr /* [gridDim.x] */) {
  // grid-stride over the K*D centre elements; per-block squared shift -> perr[block] (summed in order by the last kernel)
  float e = 0.f;
  const size_t total = (size_t)K * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / D);
    float c = 0.f, sm = 0.f;
    for (int ch = 0; ch < chunks; ++ch) { c += pcounts[(size_t)ch * K + k]; sm += psums[(size_t)ch * total + i]; }
    const float v = c > 0.f ? sm / c : 0.f;    // NaN -> 0 for empty clusters (fpk)
    new_c[i] = v;
    const float d = v - old_c[i];
    e += d * d;
  }
  __shared__ float red[8];
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    perr[blockIdx.x] = tot;
  }
}

__global__ void kmeans_err_kernel(const float* __restrict__ perr, int n, float* __restrict__ err) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { float t = 0.f; for (int i = 0; i < n; ++i) t += perr[i]; err[0] = t; }
}

}  // namespace anyloc

using namespace anyloc;

namespace anyloc {
// GEMM engines (gemm_tc.cu)
int gemm_tc_launch(const void*, const void*, int, const void*, const void*, int, int, int, int, const EpiParams&, bool,
                   cudaStream_t);
bool gemm_tc_supported(const void*, const void*, int, const void*, const void*, int, int, int, int, const EpiParams&,
                       bool);
// v3 assignment (vlad_tc.cu)
bool vlad_assign_tc_supported(const float* feats, const float* chat_tf32, int64_t R, int D, int K);
size_t vlad_assign_tc_ws_bytes(int64_t R);
int vlad_assign_tc_launch(const float* feats, const int32_t* n_valid, int n_per_img, int64_t R, int D, int K,
                          const float* chat, const float* chat_tf32, const float* cbias, const float* cnorm,
                          const float* cdnorm, int32_t* labels, float* inv_norm, cudaStream_t st, int32_t* zero_ptr = nullptr,
                          int zero_n = 0);
}  // namespace anyloc

// ANYLOC_VLAD=2 selects the v2 pipeline (coarse GEMM + full rescoring pass + shared-memory accumulate + normalise
// launch) for A/B measurements; default 3 = streaming tensor-core assignment + sorted register accumulate with the
// normalisation fused into it.
static int acc3_prefetch() {      // ANYLOC_VLAD_PREFETCH=1: accumulate3 bulk-prefetches each warp's next task into L2.  Off by
  static int v = -1;               // default: measured neutral at c2 (86.3 vs 85.7 us) and slightly negative at c5 (276 vs 270 us)
  if (v < 0) { const char* e = getenv("ANYLOC_VLAD_PREFETCH"); v = e ? atoi(e) : 0; }
  return v;
}
static int vlad_version() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ANYLOC_VLAD"); v = e ? atoi(e) : 3; }
  return v;
}

extern "C" size_t anyloc_vlad_workspace_bytes(int B, int N, int D, int K) {
  size_t R = (size_t)B * N;
  int nslices = cdiv(D, ACC_COLS);
  return 2 * align_up((size_t)K * D * 4, 256) + 2 * align_up((size_t)K * 4, 256) + align_up(R * 4, 256) * 2 +
         align_up(R * (size_t)K * 4, 256) + align_up((size_t)B * K * nslices * 4, 256) +
         align_up(((size_t)K * D + K) * 4, 256) + align_up((size_t)K * 4, 256) + vlad_assign_tc_ws_bytes((int64_t)R) + align_up((size_t)B * 4, 256) +
         4096;
}

namespace {
struct AssignBufs {
  float *chat, *chat_tf32, *cbias, *cnorm, *coarse;
  float* cdnorm = nullptr;                                                           // v3: |c^ - tf32(c^)| per centre
  int32_t* amb_count = nullptr;                                                      // reserved word of the prepared blob (round-1 work-list counter; kept zero)
  int32_t* done = nullptr; int n_done = 0;                                           // accumulate3 tickets (optional)
};

// labels (+ 1/|x|) for R rows: tensor-core coarse scores + exact rescoring when the shape allows it, else the
// FFMA kernel
int launch_assign(const float* feats, const int32_t* n_valid, int N_per_img, int64_t R, int D, int K,
                  const float* centers, int dist_mode, const AssignBufs& ab, int32_t* labels, float* inv_norm,
                  cudaStream_t st, bool prepared = false) {
  if (prepared)      // c^, tf32 copy, bias, norms and a zero work-list counter already sit in ab (anyloc_vlad_prepare)
    return vlad_assign_tc_launch(feats, n_valid, N_per_img, R, D, K, ab.chat, ab.chat_tf32, ab.cbias, ab.cnorm, ab.cdnorm,
                                 labels, inv_norm, st, ab.done, ab.n_done);
  vlad_centre_prep_kernel<<<K, 256, 0, st>>>(centers, K, D, dist_mode, ab.chat, ab.cbias, ab.chat_tf32, ab.cnorm,
                                             ab.cdnorm, ab.amb_count, ab.amb_count ? 1 : 0, ab.done, ab.done ? ab.n_done : 0);
  ANYLOC_CHECK_LAUNCH();
  if (vlad_version() >= 3 && ab.amb_count && ab.cdnorm && vlad_assign_tc_supported(feats, ab.chat_tf32, R, D, K))
    return vlad_assign_tc_launch(feats, n_valid, N_per_img, R, D, K, ab.chat, ab.chat_tf32, ab.cbias, ab.cnorm, ab.cdnorm, labels,
                                 inv_norm, st);
  EpiParams ep{ANYLOC_EPI_BIAS, ab.cbias, nullptr, nullptr, ab.coarse, nullptr, K};
  const bool fast = ab.coarse != nullptr && D <= 2048 && R >= 256 && R < (1ll << 31) &&
                    gemm_tc_supported(feats, nullptr, D, ab.chat_tf32, nullptr, D, (int)R, K, D, ep, false);
  if (fast) {
    int rc = gemm_tc_launch(feats, nullptr, D, ab.chat_tf32, nullptr, D, (int)R, K, D, ep, false, st);
    if (rc) return rc;
    const int blocks = (int)((R + 7) / 8);
    if (D <= 512)
      vlad_rescore_kernel<4><<<blocks, 256, 0, st>>>(feats, n_valid, N_per_img, R, D, K, ab.chat, ab.cbias, ab.cnorm,
                                                     ab.coarse, labels, inv_norm);
    else if (D <= 1024)
      vlad_rescore_kernel<8><<<blocks, 256, 0, st>>>(feats, n_valid, N_per_img, R, D, K, ab.chat, ab.cbias, ab.cnorm,
                                                     ab.coarse, labels, inv_norm);
    else
      vlad_rescore_kernel<16><<<blocks, 256, 0, st>>>(feats, n_valid, N_per_img, R, D, K, ab.chat, ab.cbias, ab.cnorm,
                                                      ab.coarse, labels, inv_norm);
    ANYLOC_CHECK_LAUNCH();
    return ANYLOC_OK;
  }
  int sms = device_sm_count();
  int64_t warps_needed = (R + 1) / 2;
  int blocks = (int)std::min<int64_t>((warps_needed + 7) / 8, (int64_t)sms * 8);
  if (blocks < 1) blocks = 1;
  vlad_assign_kernel<2><<<blocks, 256, 0, st>>>(feats, n_valid, N_per_img, R, D, K, ab.chat, ab.cbias, labels,
                                               inv_norm);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

bool take_assign_bufs(Workspace& w, int64_t R, int D, int K, AssignBufs* ab) {
  ab->chat = w.take<float>((size_t)K * D);
  ab->chat_tf32 = w.take<float>((size_t)K * D);
  ab->cbias = w.take<float>(K);
  ab->cnorm = w.take<float>(K);
  ab->coarse = w.take<float>((size_t)R * K);        // may be null when the caller's workspace is the small one
  ab->cdnorm = w.take<float>(K);
  ab->amb_count = w.take<int32_t>(64);
  return ab->chat && ab->chat_tf32 && ab->cbias && ab->cnorm;
}
}  // namespace

extern "C" int anyloc_vlad_assign(const float* feats, const float* centers, int R, int D, int K,
                                  int dist_mode, int32_t* labels, void* ws, size_t ws_bytes,
                                  void* stream) {
  ANYLOC_REQUIRE(feats && centers && labels && ws, "vlad_assign: null pointer");
  ANYLOC_REQUIRE(R >= 0 && D > 0 && K > 0 && D % 4 == 0, "vlad_assign: bad dims R=%d D=%d K=%d", R, D, K);
  if (R == 0) return ANYLOC_OK;
  Workspace w(ws, ws_bytes);
  AssignBufs ab;
  if (!take_assign_bufs(w, R, D, K, &ab)) { set_error("vlad_assign: workspace too small"); return ANYLOC_ERR_WORKSPACE; }
  return launch_assign(feats, nullptr, R, R, D, K, centers, dist_mode, ab, labels, nullptr, (cudaStream_t)stream);
}

// Prepared vocabulary blob (anyloc_vlad_prepare): c^ [K,D] | tf32(c^) [K,D] | bias [K] | |c^| [K] | |c^ - tf32(c^)| [K] |
// work-list counter.  Everything the per-call centre-prep launch would produce.
struct PreparedView { float *chat, *chat_tf32, *cbias, *cnorm, *cdnorm; int32_t* amb_count; };
static bool carve_prepared(void* blob, size_t bytes, int D, int K, PreparedView* pv) {
  Workspace w(blob, bytes);
  pv->chat = w.take<float>((size_t)K * D);
  pv->chat_tf32 = w.take<float>((size_t)K * D);
  pv->cbias = w.take<float>(K);
  pv->cnorm = w.take<float>(K);
  pv->cdnorm = w.take<float>(K);
  pv->amb_count = w.take<int32_t>(64);
  return pv->amb_count != nullptr;
}

extern "C" size_t anyloc_vlad_prepared_bytes(int D, int K) {
  return 2 * align_up((size_t)K * D * 4, 256) + 3 * align_up((size_t)K * 4, 256) + 256;
}

extern "C" int anyloc_vlad_prepare(const float* centers, int D, int K, int dist_mode, void* prepared,
                                   size_t prepared_bytes, void* stream) {
  ANYLOC_REQUIRE(centers && prepared, "vlad_prepare: null pointer");
  ANYLOC_REQUIRE(D > 0 && K > 0 && D % 4 == 0, "vlad_prepare: bad dims D=%d K=%d", D, K);
  ANYLOC_REQUIRE(dist_mode == ANYLOC_DIST_COSINE || dist_mode == ANYLOC_DIST_EUCLIDEAN,
                 "vlad_prepare: unknown dist_mode %d", dist_mode);
  PreparedView pv;
  if (!carve_prepared(prepared, prepared_bytes, D, K, &pv)) { set_error("vlad_prepare: blob too small"); return ANYLOC_ERR_WORKSPACE; }
  vlad_centre_prep_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(centers, K, D, dist_mode, pv.chat, pv.cbias, pv.chat_tf32,
                                                               pv.cnorm, pv.cdnorm, pv.amb_count, 1, nullptr, 0);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

static int vlad_generate_impl(const float* feats, const int32_t* n_valid, const float* centers, void* prepared,
                              size_t prepared_bytes, int B, int N, int D, int K, int dist_mode, int norm_descs,
                              int intra_norm, float* vlad, int32_t* labels_out, void* ws, size_t ws_bytes, void* stream) {
  ANYLOC_REQUIRE(feats && centers && vlad && ws, "vlad_generate: null pointer");
  ANYLOC_REQUIRE(B >= 0 && N >= 0 && D > 0 && K > 0, "vlad_generate: bad dims");
  ANYLOC_REQUIRE(D % 4 == 0, "vlad_generate: D=%d must be a multiple of 4", D);
  ANYLOC_REQUIRE(dist_mode == ANYLOC_DIST_COSINE || dist_mode == ANYLOC_DIST_EUCLIDEAN,
                 "vlad_generate: unknown dist_mode %d", dist_mode);
  if (B == 0) return ANYLOC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) { ANYLOC_CHECK_CUDA(cudaMemsetAsync(vlad, 0, (size_t)B * K * D * 4, st)); return ANYLOC_OK; }
  Workspace w(ws, ws_bytes);
  const size_t R = (size_t)B * N;
  const int nslices = cdiv(D, ACC_COLS);
  int32_t* labels = w.take<int32_t>(R);
  float* inv_norm = w.take<float>(R);
  float* partial = w.take<float>((size_t)B * K * nslices);
  AssignBufs ab;
  if (!labels || !inv_norm || !partial || !take_assign_bufs(w, (int64_t)R, D, K, &ab)) {
    set_error("vlad_generate: workspace too small (%zu bytes given)", ws_bytes);
    return ANYLOC_ERR_WORKSPACE;
  }
  const size_t smem3 = acc3_smem_bytes(N, K);
  const bool acc3 = vlad_version() >= 3 && smem3 <= 100 * 1024 && (int64_t)N * D < (1ll << 31);
  if (acc3) { ab.done = w.take<int32_t>((size_t)B); ab.n_done = B; }
  // prepared vocabulary: usable when this call takes the v3 assignment + accumulate3 route
  PreparedView pv;
  const bool use_prep = prepared && acc3 && ab.done && vlad_version() >= 3 &&
                        carve_prepared(prepared, prepared_bytes, D, K, &pv) &&
                        vlad_assign_tc_supported(feats, pv.chat_tf32, (int64_t)R, D, K);
  if (use_prep) {
    ab.chat = pv.chat; ab.chat_tf32 = pv.chat_tf32; ab.cbias = pv.cbias; ab.cnorm = pv.cnorm; ab.cdnorm = pv.cdnorm;
    ab.amb_count = pv.amb_count;
  }
  ProfScope ps(PC_VLAD, st, 4.0 * ((double)B * N * D + (double)B * K * D + (double)K * D));
  int rc = launch_assign(feats, n_valid, N, (int64_t)R, D, K, centers, dist_mode, ab, labels, inv_norm, st, use_prep);
  if (rc) return rc;
  if (acc3 && ab.done) {
    static unsigned long long attr_seen = 0;
    if (first_use_on_this_device(&attr_seen)) {
      ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(vlad_accumulate3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    // all CTAs co-resident -> the slice-CTAs of an image may wait for each other (distributed normalisation);
    // otherwise the image's last CTA normalises alone.  ANYLOC_VLAD_WAIT=0 forces the latter (A/B).
    int occ = 0;
    ANYLOC_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, vlad_accumulate3_kernel, ACC3_WARPS * 32, smem3));
    static int wait_env = -1;
    if (wait_env < 0) { const char* e = getenv("ANYLOC_VLAD_WAIT"); wait_env = e ? atoi(e) : 1; }
    const int wait_all = (wait_env && (long long)nslices * B <= (long long)occ * device_sm_count()) ? 1 : 0;
    static int timeline = -1;          // ANYLOC_VLAD_TIMELINE=1 (tools only): per-CTA phase stamps, summary on stderr
    if (timeline < 0) { const char* e = getenv("ANYLOC_VLAD_TIMELINE"); timeline = e ? atoi(e) : 0; }
    unsigned long long* dbg = nullptr;
    const size_t nctas = (size_t)nslices * B;
    if (timeline) { ANYLOC_CHECK_CUDA(cudaMalloc(&dbg, nctas * 64)); ANYLOC_CHECK_CUDA(cudaMemsetAsync(dbg, 0, nctas * 64, st)); }
    vlad_accumulate3_kernel<<<dim3(nslices, B), ACC3_WARPS * 32, smem3, st>>>(feats, labels, inv_norm, centers, N, D, K,
                                                                             norm_descs, intra_norm, vlad, partial, ab.done,
                                                                             use_prep ? ab.amb_count : nullptr, acc3_prefetch(), wait_all, dbg);
    ANYLOC_CHECK_LAUNCH();
    if (timeline) {
      std::vector<unsigned long long> h(nctas * 8);
      ANYLOC_CHECK_CUDA(cudaStreamSynchronize(st));
      ANYLOC_CHECK_CUDA(cudaMemcpy(h.data(), dbg, nctas * 64, cudaMemcpyDeviceToHost));
      cudaFree(dbg);
      unsigned long long t0 = ~0ull, t_end = 0;
      for (size_t c = 0; c < nctas; ++c) t0 = std::min(t0, h[c * 8]);
      double sum[6] = {0}, mx[6] = {0}; size_t nlast = 0;
      for (size_t c = 0; c < nctas; ++c) {
        for (int i = 0; i < 5; ++i) { double v = (double)(h[c * 8 + i] - t0) * 1e-3; sum[i] += v; mx[i] = std::max(mx[i], v); }
        if (h[c * 8 + 5]) { double v = (double)(h[c * 8 + 5] - t0) * 1e-3; sum[5] += v; mx[5] = std::max(mx[5], v); ++nlast; }
        t_end = std::max(t_end, std::max(h[c * 8 + 4], h[c * 8 + 5]));
      }
      fprintf(stderr, "[accumulate3 timeline, us since first CTA start; mean / max over %zu CTAs] start %.1f/%.1f  labels-loaded %.1f/%.1f  "
              "sorted %.1f/%.1f  tasks-done %.1f/%.1f  ticket %.1f/%.1f  normalised(%zu CTAs) %.1f/%.1f  kernel-end %.1f\n",
              nctas, sum[0] / nctas, mx[0], sum[1] / nctas, mx[1], sum[2] / nctas, mx[2], sum[3] / nctas, mx[3], sum[4] / nctas, mx[4],
              nlast, nlast ? sum[5] / nlast : 0.0, mx[5], (double)(t_end - t0) * 1e-3);
    }
    if (labels_out)
      ANYLOC_CHECK_CUDA(cudaMemcpyAsync(labels_out, labels, R * 4, cudaMemcpyDeviceToDevice, st));
    return ANYLOC_OK;
  }
  // v2: accumulate with as many row-splitting warps as shared memory allows ((1 + warps) * K * 128 floats), at most
  // 4 row-splitting warps when two CTAs then fit per SM, else what fits in one
  int warps = (int)std::min<size_t>(4, (200 * 1024) / ((size_t)K * 128 * 4) - 1);
  if (warps >= 1) {
    size_t smem = (size_t)(1 + warps) * K * 128 * 4;
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(vlad_accumulate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vlad_accumulate2_kernel<<<dim3(nslices, B), 128, smem, st>>>(feats, labels, inv_norm, centers, N, D, K, norm_descs,
                                                                 warps, vlad, partial);
  } else {
    size_t smem = ((size_t)2 * K * ACC_COLS + 2 * (size_t)N) * 4;
    ANYLOC_REQUIRE(smem <= 220 * 1024, "vlad_generate: K=%d N=%d needs %zu B shared memory", K, N, smem);
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(vlad_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vlad_accumulate_kernel<<<dim3(nslices, B), ACC_COLS, smem, st>>>(feats, labels, inv_norm, centers, N, D, K,
                                                                    norm_descs, vlad, partial);
  }
  ANYLOC_CHECK_LAUNCH();
  int ysplit = std::max(1, std::min(64, (int)(((size_t)K * D + 256 * 16 - 1) / (256 * 16))));
  vlad_normalize_kernel<<<dim3(B, ysplit), 256, 2 * K * sizeof(float), st>>>(vlad, partial, D, K, nslices,
                                                                           intra_norm);
  ANYLOC_CHECK_LAUNCH();
  if (labels_out)
    ANYLOC_CHECK_CUDA(cudaMemcpyAsync(labels_out, labels, R * 4, cudaMemcpyDeviceToDevice, st));
  return ANYLOC_OK;
}

extern "C" int anyloc_vlad_generate(const float* feats, const int32_t* n_valid, const float* centers,
                                    int B, int N, int D, int K, int dist_mode, int norm_descs,
                                    int intra_norm, float* vlad, int32_t* labels_out, void* ws,
                                    size_t ws_bytes, void* stream) {
  return vlad_generate_impl(feats, n_valid, centers, nullptr, 0, B, N, D, K, dist_mode, norm_descs, intra_norm, vlad,
                            labels_out, ws, ws_bytes, stream);
}

extern "C" int anyloc_vlad_generate_prepared(const float* feats, const int32_t* n_valid, const float* centers,
                                             void* prepared, size_t prepared_bytes, int B, int N, int D, int K,
                                             int dist_mode, int norm_descs, int intra_norm, float* vlad,
                                             int32_t* labels_out, void* ws, size_t ws_bytes, void* stream) {
  ANYLOC_REQUIRE(prepared, "vlad_generate_prepared: null prepared blob");
  return vlad_generate_impl(feats, n_valid, centers, prepared, prepared_bytes, B, N, D, K, dist_mode, norm_descs,
                            intra_norm, vlad, labels_out, ws, ws_bytes, stream);
}

// Centres for F.cosine_similarity: c / max(|c|, 1e-8)
namespace anyloc {
__global__ void vlad_soft_centre_prep_kernel(const float* __restrict__ c, int K, int D, float* __restrict__ chat) {
  int k = blockIdx.x;
  const float* row = c + (size_t)k * D;
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) { float v = row[d]; ss += v * v; }
  __shared__ float red[32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float den = fmaxf(sqrtf(red[0]), 1e-8f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) chat[(size_t)k * D + d] = row[d] / den;
}
}  // namespace anyloc

extern "C" int anyloc_vlad_generate_soft(const float* feats, const int32_t* n_valid, const float* centers,
                                         int B, int N, int D, int K, float soft_temp, int norm_descs,
                                         int intra_norm, float* vlad, float* assign_out, void* ws,
                                         size_t ws_bytes, void* stream) {
  ANYLOC_REQUIRE(feats && centers && vlad && ws, "vlad_generate_soft: null pointer");
  ANYLOC_REQUIRE(B >= 0 && N >= 0 && D > 0 && K > 0, "vlad_generate_soft: bad dims");
  ANYLOC_REQUIRE(D % 4 == 0, "vlad_generate_soft: D=%d must be a multiple of 4", D);
  ANYLOC_REQUIRE(K <= 2048, "vlad_generate_soft: K=%d > 2048", K);
  if (B == 0) return ANYLOC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) { ANYLOC_CHECK_CUDA(cudaMemsetAsync(vlad, 0, (size_t)B * K * D * 4, st)); return ANYLOC_OK; }
  Workspace w(ws, ws_bytes);
  const size_t R = (size_t)B * N;
  const int nslices = cdiv(D, ACC_COLS);
  float* inv_norm = w.take<float>(R);
  float* partial = w.take<float>((size_t)B * K * nslices);
  float* chat = w.take<float>((size_t)K * D);
  float* assign = w.take<float>(R * K);
  if (!inv_norm || !partial || !chat || !assign) {
    set_error("vlad_generate_soft: workspace too small (%zu bytes given)", ws_bytes);
    return ANYLOC_ERR_WORKSPACE;
  }
  ProfScope ps(PC_VLAD, st, 4.0 * ((double)B * N * D + (double)B * K * D + (double)K * D));
  vlad_soft_centre_prep_kernel<<<K, 256, 0, st>>>(centers, K, D, chat);
  ANYLOC_CHECK_LAUNCH();
  constexpr int ROWS = 2;
  const size_t smem = (size_t)8 * ROWS * K * 4;
  ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(vlad_soft_assign_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  int blocks = (int)std::min<int64_t>(((int64_t)(R + ROWS - 1) / ROWS + 7) / 8, (int64_t)device_sm_count() * 8);
  vlad_soft_assign_kernel<ROWS><<<std::max(blocks, 1), 256, smem, st>>>(feats, n_valid, N, (int64_t)R, D, K, chat,
                                                                       soft_temp, assign, inv_norm);
  ANYLOC_CHECK_LAUNCH();
  vlad_soft_accumulate_kernel<<<dim3(nslices, B), ACC_COLS, 0, st>>>(feats, assign, inv_norm, centers, N, D, K,
                                                                    norm_descs, vlad, partial);
  ANYLOC_CHECK_LAUNCH();
  int ysplit = std::max(1, std::min(64, (int)(((size_t)K * D + 256 * 16 - 1) / (256 * 16))));
  vlad_normalize_kernel<<<dim3(B, ysplit), 256, 2 * K * sizeof(float), st>>>(vlad, partial, D, K, nslices,
                                                                           intra_norm);
  ANYLOC_CHECK_LAUNCH();
  if (assign_out)
    ANYLOC_CHECK_CUDA(cudaMemcpyAsync(assign_out, assign, R * K * 4, cudaMemcpyDeviceToDevice, st));
  return ANYLOC_OK;
}

static int kmeans_chunks(int R, int D) {
  const int nslices = cdiv(D, ACC_COLS);
  return std::max(1, std::min(std::min((int)((R + 255) / 256), 4 * device_sm_count() / std::max(1, nslices)), 64));
}
constexpr int KMEANS_FIN_BLOCKS = 64;

extern "C" size_t anyloc_kmeans_workspace_bytes(int R, int D, int K) {
  const size_t chunks = (size_t)kmeans_chunks(R, D);
  return align_up(chunks * K * D * 4, 256) + align_up(chunks * K * 4, 256) + align_up(KMEANS_FIN_BLOCKS * 4, 256) + 256;
}

extern "C" int anyloc_kmeans_update(const float* x, const int32_t* labels, const float* old_centers,
                                    int R, int D, int K, float* new_centers, float* err_out, void* ws,
                                    size_t ws_bytes, void* stream) {
  ANYLOC_REQUIRE(x && labels && old_centers && new_centers && err_out && ws, "kmeans_update: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = kmeans_chunks(R, D);
  Workspace w(ws, ws_bytes);
  float* psums = w.take<float>((size_t)chunks * K * D);
  float* pcounts = w.take<float>((size_t)chunks * K);
  float* perr = w.take<float>(KMEANS_FIN_BLOCKS);
  if (!psums || !pcounts || !perr) {
    set_error("kmeans_update: workspace too small (%zu given, %zu needed)", ws_bytes, anyloc_kmeans_workspace_bytes(R, D, K));
    return ANYLOC_ERR_WORKSPACE;
  }
  size_t smem = ((size_t)K * ACC_COLS + K) * 4;
  ANYLOC_REQUIRE(smem <= 220 * 1024, "kmeans_update: K=%d too large", K);
  ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(kmeans_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  int nslices = cdiv(D, ACC_COLS);
  kmeans_accumulate_kernel<<<dim3(nslices, chunks), ACC_COLS, smem, st>>>(x, labels, R, D, K, psums, pcounts);
  ANYLOC_CHECK_LAUNCH();
  kmeans_finalize_kernel<<<KMEANS_FIN_BLOCKS, 256, 0, st>>>(psums, pcounts, chunks, old_centers, D, K, new_centers, perr);
  ANYLOC_CHECK_LAUNCH();
  kmeans_err_kernel<<<1, 32, 0, st>>>(perr, KMEANS_FIN_BLOCKS, err_out);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

// ====================================================================================================================
// Residual tensors and the per-image cache path (utilities.py:928-972, :843-852, :864-878).  Off the hot path: the
// reference materialises x^ - c for ALL (patch, centre) pairs ([N,K,D] fp32, 104 MB per image at c2) and, with a
// cache directory, re-builds descriptors from cached residuals / labels / soft assignments without the features.
// ====================================================================================================================
namespace anyloc {

// out[q,k,:] = x_q / max(|x_q|, 1e-12) - c_k (F.normalize when norm_descs, utilities.py:959-962).  CTA per row;
// HBM-bound on the N*K*D*4 bytes written.
__global__ void __launch_bounds__(256)
vlad_residuals_kernel(const float* __restrict__ x, const float* __restrict__ centers, int D, int K, int norm_descs,
                      float* __restrict__ out) {
  const size_t q = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + q * D);
  const int D4 = D >> 2;
  __shared__ float red[8];
  __shared__ float s_nrm;
  float ss = 0.f;
  if (norm_descs) {
    for (int d = threadIdx.x; d < D4; d += blockDim.x) { float4 v = __ldg(xr + d); ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; s_nrm = fmaxf(sqrtf(t), 1e-12f); }
    __syncthreads();
  }
  const float nrm = norm_descs ? s_nrm : 1.0f;
  float4* o = reinterpret_cast<float4*>(out + q * (size_t)K * D);
  for (int d = threadIdx.x; d < D4; d += blockDim.x) {
    float4 v = __ldg(xr + d);
    if (norm_descs) { v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm; }
    for (int k = 0; k < K; ++k) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(centers + (size_t)k * D) + d);
      o[(size_t)k * D4 + d] = make_float4(v.x - c.x, v.y - c.y, v.z - c.z, v.w - c.w);
    }
  }
}

// hard: V[k,col] = sum_{q: label_q = k} R[q,k,col] (utilities.py:858); CTA = (column slice, cluster), thread = column,
// rows in index order.  Unused clusters stay zero (:840).
__global__ void __launch_bounds__(ACC_COLS)
vlad_from_residuals_hard_kernel(const float* __restrict__ resid, const int32_t* __restrict__ labels, int N, int D, int K,
                                float* __restrict__ vlad, float* __restrict__ partial_ss) {
  const int slice = blockIdx.x, k = blockIdx.y, t = threadIdx.x, nslices = gridDim.x;
  const int col = slice * ACC_COLS + t;
  const bool colok = col < D;
  float acc = 0.f;
  for (int q = 0; q < N; ++q)
    if (__ldg(labels + q) == k && colok) acc += __ldg(resid + ((size_t)q * K + k) * D + col);
  if (colok) vlad[(size_t)k * D + col] = acc;
  __shared__ float red[ACC_COLS / 32];
  const float s = warp_sum(colok ? acc * acc : 0.f);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  if (t == 0) { float tot = 0.f; for (int w = 0; w < ACC_COLS / 32; ++w) tot += red[w]; partial_ss[(size_t)k * nslices + slice] = tot; }
}

// soft: V[k,col] = sum_q a[q,k] * sum_c R[q,c,col] (utilities.py:879-884: cluster k's weight on the residuals to ALL
// centres).  CTA = column slice, thread = column, 32 cluster accumulators per pass.
__global__ void __launch_bounds__(ACC_COLS)
vlad_from_residuals_soft_kernel(const float* __restrict__ resid, const float* __restrict__ assign, int N, int D, int K,
                                float* __restrict__ vlad, float* __restrict__ partial_ss) {
  const int slice = blockIdx.x, t = threadIdx.x, nslices = gridDim.x;
  const int col = slice * ACC_COLS + t;
  const bool colok = col < D;
  __shared__ float red[ACC_COLS / 32][32];
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kc = min(32, K - k0);
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    if (colok) {
      for (int q = 0; q < N; ++q) {
        float s = 0.f;
        for (int c = 0; c < K; ++c) s += __ldg(resid + ((size_t)q * K + c) * D + col);
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < kc) acc[j] = fmaf(__ldg(assign + (size_t)q * K + k0 + j), s, acc[j]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = (colok && j < kc) ? acc[j] : 0.f;
      if (colok && j < kc) vlad[(size_t)(k0 + j) * D + col] = v;
      const float sq = warp_sum(v * v);
      if ((t & 31) == 0) red[t >> 5][j] = sq;
    }
    __syncthreads();
    if (t < kc) {
      float tot = 0.f;
      for (int w = 0; w < ACC_COLS / 32; ++w) tot += red[w][t];
      partial_ss[(size_t)(k0 + t) * nslices + slice] = tot;
    }
  }
}

}  // namespace anyloc

extern "C" int anyloc_vlad_residuals(const float* feats, const float* centers, int N, int D, int K, int norm_descs,
                                     float* out, void* stream) {
  ANYLOC_REQUIRE(feats && centers && out, "vlad_residuals: null pointer");
  ANYLOC_REQUIRE(N >= 0 && D > 0 && K > 0 && D % 4 == 0, "vlad_residuals: bad dims N=%d D=%d K=%d", N, D, K);
  if (N == 0) return ANYLOC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(PC_VLAD, st, 4.0 * ((double)N * D + (double)K * D + (double)N * K * D));
  vlad_residuals_kernel<<<N, 256, 0, st>>>(feats, centers, D, K, norm_descs, out);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

extern "C" size_t anyloc_vlad_from_residuals_workspace_bytes(int D, int K) {
  return align_up((size_t)K * cdiv(D, ACC_COLS) * 4, 256) + 256;
}

extern "C" int anyloc_vlad_from_residuals(const float* resid, const int32_t* labels, const float* assign, int N, int D,
                                          int K, int intra_norm, float* vlad, void* ws, size_t ws_bytes, void* stream) {
  ANYLOC_REQUIRE(resid && vlad && ws, "vlad_from_residuals: null pointer");
  ANYLOC_REQUIRE((labels != nullptr) != (assign != nullptr), "vlad_from_residuals: pass labels (hard) OR assign (soft)");
  ANYLOC_REQUIRE(N >= 0 && D > 0 && K > 0, "vlad_from_residuals: bad dims N=%d D=%d K=%d", N, D, K);
  cudaStream_t st = (cudaStream_t)stream;
  const int nslices = cdiv(D, ACC_COLS);
  Workspace w(ws, ws_bytes);
  float* partial = w.take<float>((size_t)K * nslices);
  if (!partial) { set_error("vlad_from_residuals: workspace too small"); return ANYLOC_ERR_WORKSPACE; }
  ProfScope ps(PC_VLAD, st, 4.0 * ((double)N * K * D + (double)K * D));
  if (labels) vlad_from_residuals_hard_kernel<<<dim3(nslices, K), ACC_COLS, 0, st>>>(resid, labels, N, D, K, vlad, partial);
  else vlad_from_residuals_soft_kernel<<<nslices, ACC_COLS, 0, st>>>(resid, assign, N, D, K, vlad, partial);
  ANYLOC_CHECK_LAUNCH();
  int ysplit = std::max(1, std::min(64, (int)(((size_t)K * D + 256 * 16 - 1) / (256 * 16))));
  vlad_normalize_kernel<<<dim3(1, ysplit), 256, 2 * K * sizeof(float), st>>>(vlad, partial, D, K, nslices, intra_norm);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}
