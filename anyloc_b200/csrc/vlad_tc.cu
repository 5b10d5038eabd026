// Hard-assignment VLAD, v3 assignment stage (reference: fpk.KMeans.predict called at
// /root/reference/utilities.py:849; row norms of F.normalize at :959-960).
//
// vlad_assign_tc_kernel -- ONE streaming pass over the features that yields, per row,
//   * 1/max(|x|, 1e-12)                                    (exact fp32, from the staged shared-memory tiles)
//   * the label, when the coarse tensor-core scores decide it rigorously, else an entry in the "ambiguous" work list
// Persistent CTA per SM, tile = 128 rows x all of D, streamed as 128-byte k-blocks through a deep TMA ring
// (10-11 stages at K=32: ~170 KB of features in flight per SM, which is what a latency-bound HBM stream needs):
//   warp 0    TMA producer: X box [128 rows x 32 floats] + centre box [K rows x 32 floats] (L2-resident) per stage
//   warp 1    MMA issuer: tcgen05.mma kind::tf32 M128 x N=K x K8 straight from the raw fp32 words (the tensor core
//             truncates to tf32), coarse scores S~ accumulate over all of D in TMEM (2 x 128 columns, double buffered)
//   warp 2    TMEM allocator
//   warps 4-7 thread = row: sum of squares of every staged k-block (128B-swizzled shared-memory reads, conflict free),
//             (and of what the tensor core's tf32 truncation drops), then the tile epilogue: S~ from TMEM, candidate set
//             {k : S~_k >= max - 2 eps} with a rigorous per-row bound eps on |S~ - S| built from those MEASURED norms
//             (see the epilogue).  One candidate -> that IS the exact-fp32 argmax.
//             Several -> (row, candidate mask) goes to the tile's shared-memory list (double buffered).
//   warps 8-15 exact re-scoring of the listed rows while the stream continues: warp per row, the row re-read from L2
//             (it passed through this SM microseconds ago), exact fp32 dot products for the candidates only (the
//             arithmetic of the v2 rescoring kernel), first-max argmax -> lowest index wins exact ties, all-zero rows
//             get label 0.  (Round 1 ran this as a separate launch over a global work list: 12-16 us at c2, 30 us at c5.)
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace anyloc {
namespace vtc {
using namespace tc;

constexpr int BM = 128;
constexpr int A_BYTES = BM * 128;          // 16 KB: 128 rows x 128 B
constexpr int R_WARPS = 8;                 // exact re-scoring warps (warps 8..15)
constexpr int THREADS = 256 + R_WARPS * 32;
constexpr int TMEM_COLS = 256;             // 2 accumulators x 128 columns
constexpr int MAX_STAGES = 12;
constexpr int MAX_K = 128;
constexpr int BAR_BYTES = 320;             // (2*MAX_STAGES + 4) mbarriers, the TMEM slot, 4 list mbarriers
constexpr int VEC_BYTES = 3 * MAX_K * 4;   // cbias + cnorm + cdnorm
constexpr int LIST_BYTES = 2 * (BM * 4 + BM * (MAX_K / 32) * 4) + 16;   // 2 x {ambiguous rows, candidate masks} + counters

struct AssignParams {
  const int32_t* n_valid; int n_per_img; int R; int D; int K;
  const float* cbias; const float* cnorm; const float* cdnorm;
  int32_t* labels; float* inv_norm;
  const float* x; const float* chat;  // features [R,D] and the exact fp32 c^ [K,D]: the in-kernel re-scoring reads them
  int32_t* zero_ptr; int zero_n;     // cleared here for a LATER launch on the stream (accumulate tickets), nullable
  int stages; int stage_bytes; int n_mma; int burst;
  int tile_rows;     // rows per tile (<= BM, multiple of 8), chosen on the host so that the tiles fill whole waves of the grid
  int diag;      // timing experiments only (tools/, results invalid): bit 0 = row-norm math off, bit 1 = MMAs off
};

// Exact fp32 re-scoring of ONE ambiguous row by one warp: dot products with the candidate centres only (mask), in
// ascending k with a strict >, so the lowest index wins exact ties and all-zero rows get label 0 -- the arithmetic of the
// v2 re-scoring pass.  The row (D <= 2048) lives in registers.
__device__ __forceinline__ void rescore_row(const AssignParams& p, const float* s_cbias, int64_t row, const uint32_t* mask_w,
                                            int lane) {
  constexpr int MAXV = 16;
  const int D4 = p.D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(p.x + row * (int64_t)p.D);
  float4 v[MAXV];
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    const int d = lane + j * 32;
    v[j] = d < D4 ? __ldg(xr + d) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float best = -INFINITY; int bestk = 0;
  for (int w = 0; w < MAX_K / 32; ++w) {
    uint32_t mask = mask_w[w];
    const int k0 = w * 32;
    while (mask) {
      int kk[4]; int nc = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (mask) { kk[q] = k0 + __ffs(mask) - 1; mask &= mask - 1; ++nc; } else kk[q] = kk[0];
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < MAXV; ++j) {
        const int d = lane + j * 32;
        if (d < D4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(p.chat + (size_t)kk[q] * p.D) + d);
            acc[q] = fmaf(v[j].x, c.x, acc[q]); acc[q] = fmaf(v[j].y, c.y, acc[q]);
            acc[q] = fmaf(v[j].z, c.z, acc[q]); acc[q] = fmaf(v[j].w, c.w, acc[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sc = warp_sum(acc[q]) + s_cbias[kk[q]];
        if (q < nc && sc > best) { best = sc; bestk = kk[q]; }
      }
    }
  }
  if (lane == 0) p.labels[row] = bestk;
}

__global__ void __launch_bounds__(THREADS, 1)
vlad_assign_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_c,
                      const AssignParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* bar_area = smem + p.stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);          // [MAX_STAGES]
  uint64_t* empty_bar = full_bar + MAX_STAGES;                          // [MAX_STAGES]
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;                         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_cbias = reinterpret_cast<float*>(bar_area + BAR_BYTES);      // [MAX_K]
  float* s_cnorm = s_cbias + MAX_K;                                     // [MAX_K]
  float* s_cdnorm = s_cnorm + MAX_K;                                    // [MAX_K]
  uint64_t* list_full = tempty_bar + 2 + 1;                             // [2] epilogue -> re-scoring warps (after the TMEM slot word pair)
  uint64_t* list_free = list_full + 2;                                  // [2] re-scoring warps -> epilogue
  uint64_t* last_full = list_free + 2;                                  // [1] the CTA's LAST list is complete (single phase)
  int32_t* amb_row = reinterpret_cast<int32_t*>(bar_area + BAR_BYTES + VEC_BYTES);    // [2][BM]
  uint32_t* amb_msk = reinterpret_cast<uint32_t*>(amb_row + 2 * BM);                   // [2][BM][MAX_K/32]
  int32_t* amb_n = reinterpret_cast<int32_t*>(amb_msk + 2 * BM * (MAX_K / 32));        // [2]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_tiles = (p.R + p.tile_rows - 1) / p.tile_rows;
  const int num_k = (p.D + 31) / 32;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 5); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), 4);
      mbar_init(smem_u32(list_full + s), 4); mbar_init(smem_u32(list_free + s), 1);
      amb_n[s] = 0;
    }
    mbar_init(smem_u32(last_full), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = blockIdx.x * THREADS + threadIdx.x; i < p.zero_n; i += gridDim.x * THREADS) p.zero_ptr[i] = 0;
  for (int k = threadIdx.x; k < MAX_K; k += THREADS) {
    s_cbias[k] = k < p.K ? p.cbias[k] : 0.f;
    s_cnorm[k] = k < p.K ? p.cnorm[k] : 0.f;
    s_cdnorm[k] = k < p.K ? p.cdnorm[k] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer
      // Stages are (re)filled in bursts of `burst` consecutive k-blocks: the TMA requests of one burst hit
      // burst * 128 contiguous bytes of every row back to back, which the DRAM controller can serve from one open
      // page (a lone 128-byte access per row every few hundred ns re-opens the page each time).
      const uint32_t tx_bytes = (uint32_t)(p.tile_rows * 128 + p.n_mma * 128);   // the box holds tile_rows rows
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tile * p.tile_rows;
        for (int kb0 = 0; kb0 < num_k; kb0 += p.burst) {
          const int g = min(p.burst, num_k - kb0);
          for (int j = 0; j < g; ++j) {
            int sj = stage + j; uint32_t pj = phase;
            if (sj >= p.stages) { sj -= p.stages; pj ^= 1; }
            mbar_wait(smem_u32(empty_bar + sj), pj ^ 1);
          }
          for (int j = 0; j < g; ++j) {
            const uint32_t fb = smem_u32(full_bar + stage);
            mbar_expect_tx(fb, tx_bytes);
            const uint32_t sbase = smem_u32(smem + stage * p.stage_bytes);
            tma_load_2d(sbase, &tm_x, fb, (kb0 + j) * 32, m0);
            tma_load_2d(sbase + A_BYTES, &tm_c, fb, (kb0 + j) * 32, 0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer (tf32 in, fp32 accumulate, M128 x n_mma x K8)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * p.stage_bytes);
          const uint64_t a = make_desc(sbase), b = make_desc(sbase + A_BYTES);
          if (!(p.diag & 2)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              umma<false>(d_tmem, a + adv, b + adv, idesc, (uint32_t)((kb | k) != 0));
            }
          }
          umma_commit(smem_u32(empty_bar + stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(tfull_bar + acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------- row norms from the staged tiles + tile epilogue; thread = row
    const int q = warp & 3;                          // TMEM lane quarter == row quarter of the tile
    const int rt = q * 32 + lane;
    const int sw = rt & 7;
    float cmax = 0.f, dcmax = 0.f;
    for (int k = 0; k < p.K; ++k) { cmax = fmaxf(cmax, s_cnorm[k]); dcmax = fmaxf(dcmax, s_cdnorm[k]); }
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      // |x|^2 and |x - tf32_trunc(x)|^2 (what the tensor core drops), two partial sums each: the FMA chains of a stage
      // are half as deep, and the tile is read with ld.shared (the generic-address form costs an address translation)
      float ss0 = 0.f, ss1 = 0.f, dd0 = 0.f, dd1 = 0.f;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(smem_u32(full_bar + stage), phase);
        const uint32_t rowa = smem_u32(smem + stage * p.stage_bytes + rt * 128);
        if (!(p.diag & 1)) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(rowa + (uint32_t)((c ^ sw) << 4)));
            const float dx = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            const float dy = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            const float dz = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            const float dw = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            if (c & 1) { ss1 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; dd1 += dx * dx + dy * dy + dz * dz + dw * dw; }
            else       { ss0 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; dd0 += dx * dx + dy * dy + dz * dz + dw * dw; }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(empty_bar + stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      const float ss = ss0 + ss1, dd = dd0 + dd1;
      const float xn = sqrtf(ss);
      mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
      float smax = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < MAX_K / 16; ++cc) {
        if (cc * 16 < p.n_mma) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(cc * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = cc * 16 + j;
            if (k < p.K) smax = fmaxf(smax, v[j] + s_cbias[k]);
          }
        }
      }
      // S~_k - S_k = -sum d_i c^_i - sum x_i e_i + sum d_i e_i + (accumulation), with d = x - trunc_tf32(x) (its norm is
      // MEASURED per row above) and e = c^ - tf32(c^) (norm per centre from the prep kernel), so by Cauchy-Schwarz
      //   |S~_k - S_k| <= |d||c^| + |x||e| + acc,
      // acc <= 1e-4 |x||c^|: D/8 <= 256 tensor-core accumulation steps, each truncating the running sum; the measured
      // truncation bias of this engine is 2^-26 of the sum per MMA (-9.4e-6 over 576 MMAs), i.e.
      // <= 256 * 2^-24 = 1.5e-5 even at four times the mean, so 1e-4 leaves a 6x margin.  The 1 % on top covers the
      // fp32 evaluation of the norms and the second-order term.  ~3x tighter than the a-priori 2^-9 |x||c^| of v2
      // (|d| is typically 0.4 * 2^-10 |x|): about a third of the rows that v2 had to re-score stay ambiguous.
      const float eps = 1.01f * (sqrtf(dd) * cmax + xn * (dcmax + 1.0e-4f * cmax));
      const float thresh = smax - 2.0f * eps - 1e-30f;
      uint32_t mask[MAX_K / 32];
#pragma unroll
      for (int i = 0; i < MAX_K / 32; ++i) mask[i] = 0u;
      int cnt = 0, first = 0;
#pragma unroll
      for (int cc = 0; cc < MAX_K / 16; ++cc) {
        if (cc * 16 < p.n_mma) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(cc * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = cc * 16 + j;
            if (k < p.K && v[j] + s_cbias[k] >= thresh) {
              if (cnt == 0) first = k;
              ++cnt;
              mask[cc >> 1] |= 1u << (k & 31);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      // rows [tile_rows, BM) of the shared-memory tile are never written by TMA: their scores are garbage and ignored
      const int64_t m = (int64_t)tile * p.tile_rows + rt;
      bool amb = false;
      if (rt < p.tile_rows && m < p.R) {
        bool valid = true;
        if (p.n_valid) { const int b = (int)(m / p.n_per_img), n = (int)(m - (int64_t)b * p.n_per_img); valid = n < p.n_valid[b]; }
        if (p.inv_norm) p.inv_norm[m] = 1.0f / fmaxf(xn, 1e-12f);
        if (!valid) p.labels[m] = -1;
        else if (cnt <= 1) p.labels[m] = first;          // cnt == 0 only with NaN scores: label 0 like the exact path
        else amb = true;
      }
      // ambiguous rows -> this tile's shared-memory list (double buffered); the re-scoring warps take it from there
      const int buf = ti & 1;
      mbar_wait(smem_u32(list_free + buf), (uint32_t)(((ti >> 1) & 1) ^ 1));     // list `buf` consumed (tile ti - 2)
      if (amb) {
        const int idx = atomicAdd(&amb_n[buf], 1);                             // < BM by construction
        amb_row[buf * BM + idx] = (int32_t)m;
#pragma unroll
        for (int i = 0; i < MAX_K / 32; ++i) amb_msk[(buf * BM + idx) * (MAX_K / 32) + i] = mask[i];
      }
      __syncwarp();
      // the last list goes to ALL warps through its own single-phase barrier: warps without a role reach the tail at
      // once and have not followed the phases of list_full (a parity wait there would alias an earlier phase)
      if (lane == 0) mbar_arrive(smem_u32(tile + (int)gridDim.x >= num_tiles ? last_full : list_full + buf));
    }
  } else if (warp >= 8) {
    // ---------------------------------------- exact fp32 re-scoring of the tile's ambiguous rows (warp per row): the
    // row was streamed through this SM microseconds ago, so it is re-read from L2; same arithmetic (and therefore the
    // same labels) as the former vlad_rescore_amb_kernel: candidates in ascending k, strict >: lowest index wins
    // exact ties, all-zero rows get label 0.
    const int rw = warp - 8;
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      if (tile + (int)gridDim.x >= num_tiles) break;      // the CTA's LAST list is re-scored by all 16 warps below
      const int buf = ti & 1;
      mbar_wait(smem_u32(list_full + buf), (uint32_t)((ti >> 1) & 1));
      const int n = amb_n[buf];
      for (int i = rw; i < n; i += R_WARPS)
        rescore_row(p, s_cbias, amb_row[buf * BM + i], amb_msk + (buf * BM + i) * (MAX_K / 32), lane);
      asm volatile("bar.sync 2, %0;" ::"n"(R_WARPS * 32) : "memory");       // every re-scoring warp is done with list `buf`
      if (rw == 0 && lane == 0) { amb_n[buf] = 0; mbar_arrive(smem_u32(list_free + buf)); }
    }
  }
  // ---- the CTA's last tile: nothing is left to overlap with, so ALL 16 warps re-score its list (at c2 every CTA has a
  // single tile: 13-14 ambiguous rows per tile, one round instead of two behind eight warps)
  {
    const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (my_tiles > 0) {
      const int tl = my_tiles - 1, buf = tl & 1;
      mbar_wait(smem_u32(last_full), 0u);
      const int n = amb_n[buf];
      for (int i = warp; i < n; i += THREADS / 32)
        rescore_row(p, s_cbias, amb_row[buf * BM + i], amb_msk + (buf * BM + i) * (MAX_K / 32), lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace vtc

// Shapes the tensor-core assignment handles; everything else stays on the v2 / FFMA kernels.
bool vlad_assign_tc_supported(const float* feats, const float* chat_tf32, int64_t R, int D, int K) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return R >= 256 && R < (1ll << 31) - 256 && D >= 32 && D <= 2048 && (D % 4) == 0 && K >= 1 && K <= vtc::MAX_K &&
         al16(feats) && al16(chat_tf32);
}

size_t vlad_assign_tc_ws_bytes(int64_t R) {
  (void)R;           // the ambiguous-row lists live in shared memory since the re-scoring moved into the kernel
  return 256;
}

// feats [R,D]; chat (exact fp32 c^), chat_tf32 (rounded copy), cbias / cnorm [K] come from vlad_centre_prep_kernel;
// labels [R], inv_norm [R] (nullable).
int vlad_assign_tc_launch(const float* feats, const int32_t* n_valid, int n_per_img, int64_t R, int D, int K,
                          const float* chat, const float* chat_tf32, const float* cbias, const float* cnorm,
                          const float* cdnorm, int32_t* labels, float* inv_norm, cudaStream_t st, int32_t* zero_ptr,
                          int zero_n) {
  using namespace vtc;
  CUtensorMap mx, mc;
  int rc;
  const int n_mma = (K + 15) / 16 * 16;
  // Rows per tile: the largest multiple of 8 <= 128 that minimises (waves of the persistent grid) x (rows per tile), so
  // that the tiles fill whole waves (c2: 16 928 rows -> 146 tiles of 116 rows on 148 SMs instead of 133 of 128).
  const int sms = device_sm_count();
  int tile_rows = BM;
  {
    long long best = -1;
    for (int tr = BM; tr >= 64; tr -= 8) {
      const long long tiles_tr = (R + tr - 1) / tr, waves = (tiles_tr + sms - 1) / sms, cost = waves * tr;
      if (best < 0 || cost < best) { best = cost; tile_rows = tr; }
    }
  }
  if ((rc = tc::make_map(&mx, feats, (int)R, D, D, tile_rows, false))) return rc;
  if ((rc = tc::make_map(&mc, chat_tf32, K, D, D, n_mma, false))) return rc;
  AssignParams p;
  p.n_valid = n_valid; p.n_per_img = n_per_img; p.R = (int)R; p.D = D; p.K = K;
  p.cbias = cbias; p.cnorm = cnorm; p.cdnorm = cdnorm; p.labels = labels; p.inv_norm = inv_norm;
  p.x = feats; p.chat = chat; p.tile_rows = tile_rows;
  p.zero_ptr = zero_ptr; p.zero_n = zero_ptr ? zero_n : 0;
  p.n_mma = n_mma;
  p.stage_bytes = A_BYTES + n_mma * 128;
  int dev = 0, max_smem = 0;
  ANYLOC_CHECK_CUDA(cudaGetDevice(&dev));
  ANYLOC_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  static unsigned long long attr_seen = 0;
  if (first_use_on_this_device(&attr_seen))
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(vlad_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  const int fixed = 1024 + BAR_BYTES + VEC_BYTES + LIST_BYTES;
  p.stages = std::min(MAX_STAGES, (max_smem - fixed) / p.stage_bytes);
  const int num_k = (D + 31) / 32;
  p.stages = std::max(2, std::min(p.stages, std::max(2, num_k)));
  static int burst_env = -1;             // ANYLOC_VLAD_TMA_BURST: k-blocks issued back to back by the producer (A/B knob)
  if (burst_env < 0) { const char* e = getenv("ANYLOC_VLAD_TMA_BURST"); burst_env = e ? atoi(e) : 4; }
  static int stages_env = -1, diag_env = -1;   // A/B and timing-experiment knobs (tools/ only)
  if (stages_env < 0) { const char* e = getenv("ANYLOC_VLAD_STAGES"); stages_env = e ? atoi(e) : 0; }
  if (diag_env < 0) { const char* e = getenv("ANYLOC_VLAD_DIAG"); diag_env = e ? atoi(e) : 0; }
  if (stages_env >= 2) p.stages = std::min(p.stages, stages_env);
  p.diag = diag_env;
  p.burst = std::max(1, std::min(burst_env, p.stages / 2));
  const size_t smem = (size_t)p.stages * p.stage_bytes + fixed;
  const int tiles = (int)((R + tile_rows - 1) / tile_rows);
  const int grid = std::min(tiles, device_sm_count());
  vlad_assign_tc_kernel<<<grid, THREADS, smem, st>>>(mx, mc, p);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

}  // namespace anyloc
