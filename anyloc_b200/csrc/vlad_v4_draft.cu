// =====================================================================================================================
// DRAFT -- NOT part of libanyloc_b200.so (anyloc_b200/build.py skips *_draft.cu) and NEVER RUN ON A GPU YET.
// It only has to compile (nvcc -c).  It is the persistent two-group VLAD kernel planned in DESIGN.md section 8.1,
// written while no GPU time was left in round 1 so that the next round starts from code instead of a sketch; every
// claim about it is a plan, not a measurement.  The product path is vlad_tc.cu + vlad.cu (v3).
// =====================================================================================================================
// Hard-assignment VLAD (reference: /root/reference/utilities.py:819-926; assignment fpk.KMeans.predict :849; residuals
// :956-962) in ONE launch with ONE pass over the features in DRAM:
//
//   CTA per SM, 16 warps.
//   S-side (warps 0-7) = the v3 assignment pipeline (vlad_tc.cu) per 128-row tile, static tile striding: TMA ring of
//     128-byte k-blocks, tcgen05 tf32 coarse scores in TMEM, row norms from the staged tiles, tile epilogue with the
//     measured bound.  Decided labels go to global memory, AMBIGUOUS rows to a double-buffered shared-memory list.
//   W-side (warps 8-15) = eight worker warps driven by a tiny priority scheduler (their leader picks a job, a named
//     barrier publishes it):
//       RESCORE(i)  exact fp32 re-scoring of the CTA's i-th tile list (rows re-read from L2), then the tile's rows are
//                   credited to rows_done[image]; the CTA that completes an image queues a SORT for it
//       SORT(b)     stable counting sort of image b's rows by label, ONCE per image, into global tables (row offsets,
//                   1/|x| in sorted order, cluster / task / slot offsets); then image_ready[b] = 1 (release)
//       ACC(b, s)   accumulate item (image, 128-column slice), claimed in image order with a CAS and only when
//                   image_ready[b] is already set, so no job ever waits inside: the v3 task loop (8 x 512-byte row
//                   segments in flight per warp, sums in registers, multi-task clusters combined from shared-memory
//                   slots in task order => one fixed summation order), ticket, the image's last slice normalises
//       EXIT        own tiles handled and every accumulate item claimed
//   No job blocks on another job, S-side only ever waits for its own CTA's W-side (list buffer free), W-side RESCORE
//   has priority, so the dependency graph is acyclic: tile lists -> rows_done -> sort -> ready -> accumulate items.
//
// Known gaps to close when this first meets a GPU (besides whatever the first run shows):
//   * registers: 512 threads x 128 leave the re-scoring row (MAXV = 16) spilling; give the W warps more and the S warps
//     fewer with setmaxnreg, as gemm_tc.cu does;
//   * sort_q holds 8 images: enough for N >= 32 with SORT drained between RESCOREs, not proven in general;
//   * accumulate items are claimed strictly in image order (head-of-line blocking if images complete out of order);
//   * the image's last slice normalises the whole descriptor alone (v3 spreads that over the slices, vlad.cu).
//
// Expected: accumulate items read the image from L2 (it was streamed microseconds earlier) while the
// tile stream keeps DRAM busy; at c5 (4.6 tiles per CTA) the two overlap.
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace anyloc {
namespace v4 {
using namespace tc;

constexpr int BM = 128;
constexpr int A_BYTES = BM * 128;
constexpr int MAX_K = 128;
constexpr int MAX_STAGES = 10;
constexpr int S_WARPS = 8, W_WARPS = 8;
constexpr int THREADS = (S_WARPS + W_WARPS) * 32;       // 512
constexpr int W_THREADS = W_WARPS * 32;
constexpr int TMEM_COLS = 256;
constexpr int SEG = 64;                                   // rows per accumulate task
constexpr int BAR_BYTES = 320;                            // (2*MAX_STAGES + 4 + 4) mbarriers + TMEM slot
constexpr int VEC_BYTES = 3 * MAX_K * 4;                  // cbias, cnorm, cdnorm
constexpr int LIST_BYTES = 2 * BM * 4 + 2 * BM * (MAX_K / 32) * 4 + 64;   // ambiguous rows + masks (x2) + counters
constexpr int SORTQ = 8;

enum Job { J_IDLE = 0, J_RESCORE, J_SORT, J_ACC, J_EXIT };

struct Params {
  const float* x; const int32_t* n_valid; int B, N, D, K; int R;
  const float* centers;                                   // raw [K,D] (residuals)
  const float *chat, *cbias, *cnorm, *cdnorm;             // prepared vocabulary (vlad_centre_prep_kernel)
  int norm_descs, intra_norm;
  int32_t* labels; float* inv_norm; float* vlad; float* partial_ss;
  // scheduler state, ONE contiguous block zeroed by a memset before the launch
  int32_t* next_acc; int32_t* rows_done; int32_t* image_ready; int32_t* done;
  // per-image tables written by SORT, read by ACC
  int32_t* t_ooff; float* t_inv; int32_t* t_start; int32_t* t_tstart; int32_t* t_sbase; int32_t* t_taskk;
  int stages, stage_bytes, n_mma, nslices, max_tasks, max_slots, num_tiles;
};

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {       // one non-blocking probe
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void w_sync() { asm volatile("bar.sync 2, %0;" ::"n"(W_THREADS) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
template <int MAXV>      // float4 per lane of a row held in registers while it is re-scored: D <= 128 * MAXV
__global__ void __launch_bounds__(THREADS, 1)
vlad_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* bar_area = smem + p.stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);          // [MAX_STAGES]
  uint64_t* empty_bar = full_bar + MAX_STAGES;                          // [MAX_STAGES]
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;                         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                                 // [2]
  uint64_t* list_full = tempty_bar + 2;                                 // [2] S -> W: tile list complete (4 epilogue warps)
  uint64_t* list_free = list_full + 2;                                  // [2] W -> S: list consumed (1 arrival)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(list_free + 2);
  float* s_cbias = reinterpret_cast<float*>(bar_area + BAR_BYTES);
  float* s_cnorm = s_cbias + MAX_K;
  float* s_cdnorm = s_cnorm + MAX_K;
  int32_t* amb_row = reinterpret_cast<int32_t*>(bar_area + BAR_BYTES + VEC_BYTES);        // [2][BM]
  uint32_t* amb_mask = reinterpret_cast<uint32_t*>(amb_row + 2 * BM);                     // [2][BM][MAX_K/32]
  int32_t* amb_n = reinterpret_cast<int32_t*>(amb_mask + 2 * BM * (MAX_K / 32));          // [2]
  int32_t* s_job = amb_n + 2;                                                             // [2] job, argument
  int32_t* s_next_task = s_job + 2;
  int32_t* s_last = s_next_task + 1;
  float* s_gnorm = reinterpret_cast<float*>(s_last + 1);
  int32_t* sort_q = reinterpret_cast<int32_t*>(s_gnorm + 1);                              // [SORTQ]
  int32_t* arena = reinterpret_cast<int32_t*>(bar_area + BAR_BYTES + VEC_BYTES + LIST_BYTES + 64);   // W-side job memory

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_k = (p.D + 31) / 32;
  const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 5); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), 4);
      mbar_init(smem_u32(list_full + s), 4); mbar_init(smem_u32(list_free + s), 1);
    }
    amb_n[0] = amb_n[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
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
      // ------------------------------------------------ S: TMA producer
      const uint32_t tx_bytes = (uint32_t)(A_BYTES + p.n_mma * 128);
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t fb = smem_u32(full_bar + stage);
          mbar_expect_tx(fb, tx_bytes);
          const uint32_t sbase = smem_u32(smem + stage * p.stage_bytes);
          tma_load_2d(sbase, &tm_x, fb, kb * 32, m0);
          tma_load_2d(sbase + A_BYTES, &tm_c, fb, kb * 32, 0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ S: MMA issuer
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * p.stage_bytes);
          const uint64_t a = make_desc(sbase), b = make_desc(sbase + A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma<false>(d_tmem, a + adv, b + adv, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(smem_u32(empty_bar + stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(tfull_bar + acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < S_WARPS) {
    // ---------------------------------------- S: row norms + tile epilogue; thread = row (as vlad_tc.cu)
    const int q = warp & 3, rt = q * 32 + lane, sw = rt & 7;
    float cmax = 0.f, dcmax = 0.f;
    for (int k = 0; k < p.K; ++k) { cmax = fmaxf(cmax, s_cnorm[k]); dcmax = fmaxf(dcmax, s_cdnorm[k]); }
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int lp = i & 1;                                   // list buffer of this tile
      float ss0 = 0.f, ss1 = 0.f, dd0 = 0.f, dd1 = 0.f;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(smem_u32(full_bar + stage), phase);
        const uint32_t rowa = smem_u32(smem + stage * p.stage_bytes + rt * 128);
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
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(empty_bar + stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      const float ss = ss0 + ss1, dd = dd0 + dd1, xn = sqrtf(ss);
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
          for (int j = 0; j < 16; ++j) { const int k = cc * 16 + j; if (k < p.K) smax = fmaxf(smax, v[j] + s_cbias[k]); }
        }
      }
      const float eps = 1.01f * (sqrtf(dd) * cmax + xn * (dcmax + 1.0e-4f * cmax));      // vlad_tc.cu, tile epilogue
      const float thresh = smax - 2.0f * eps - 1e-30f;
      uint32_t mask[MAX_K / 32];
#pragma unroll
      for (int j = 0; j < MAX_K / 32; ++j) mask[j] = 0u;
      int cnt = 0, first = 0;
#pragma unroll
      for (int cc = 0; cc < MAX_K / 16; ++cc) {
        if (cc * 16 < p.n_mma) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(cc * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = cc * 16 + j;
            if (k < p.K && v[j] + s_cbias[k] >= thresh) { if (cnt == 0) first = k; ++cnt; mask[cc >> 1] |= 1u << (k & 31); }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      // the list buffer of tile i-2 must have been consumed by the workers
      if (i >= 2) mbar_wait(smem_u32(list_free + lp), (uint32_t)(((i >> 1) - 1) & 1));
      const int64_t m = (int64_t)tile * BM + rt;
      if (m < p.R) {
        bool valid = true;
        if (p.n_valid) { const int b = (int)(m / p.N), n = (int)(m - (int64_t)b * p.N); valid = n < p.n_valid[b]; }
        p.inv_norm[m] = 1.0f / fmaxf(xn, 1e-12f);
        if (!valid) p.labels[m] = -1;
        else if (cnt <= 1) p.labels[m] = first;
        else {
          const int idx = atomicAdd(&amb_n[lp], 1);           // < BM by construction
          amb_row[lp * BM + idx] = (int32_t)m;
#pragma unroll
          for (int j = 0; j < MAX_K / 32; ++j) amb_mask[(lp * BM + idx) * (MAX_K / 32) + j] = mask[j];
        }
      }
      __threadfence();                                        // labels / inv_norm of this tile before the hand-over
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(list_full + lp));
    }
  } else if (warp >= S_WARPS) {
    // =================================================================================================== W-side
    const int wt = (int)threadIdx.x - S_WARPS * 32;           // 0..255
    const int ww = wt >> 5;                                   // worker warp 0..7
    const int n_items = p.B * p.nslices;
    int handled = 0, sq_head = 0, sq_tail = 0;                // leader-only state (thread wt == 0)
    const int D4 = p.D >> 2;
    for (;;) {
      if (wt == 0) {
        int job = J_IDLE, arg = 0;
        if (handled < my_tiles && mbar_test(smem_u32(list_full + (handled & 1)), (uint32_t)((handled >> 1) & 1))) {
          job = J_RESCORE; arg = handled;
        } else if (sq_head != sq_tail) {
          job = J_SORT; arg = sort_q[sq_head & (SORTQ - 1)]; ++sq_head;
        } else {
          const int it = ld_acquire(p.next_acc);
          if (it < n_items) {
            if (ld_acquire(p.image_ready + it / p.nslices) && atomicCAS(p.next_acc, it, it + 1) == it) { job = J_ACC; arg = it; }
          } else if (handled == my_tiles) {
            job = J_EXIT;
          }
        }
        if (job == J_IDLE) __nanosleep(200);
        s_job[0] = job; s_job[1] = arg;
      }
      w_sync();
      const int job = s_job[0], arg = s_job[1];
      if (job == J_EXIT) break;

      if (job == J_RESCORE) {
        // ------------------------------------------------------------------ exact re-scoring of the i-th tile list
        const int lp = arg & 1;
        const int n = amb_n[lp];
        for (int r = ww; r < n; r += W_WARPS) {
          const int64_t row = amb_row[lp * BM + r];
          const float4* xr = reinterpret_cast<const float4*>(p.x + row * (int64_t)p.D);
          float4 v[MAXV];
#pragma unroll
          for (int j = 0; j < MAXV; ++j) { const int d = lane + j * 32; v[j] = d < D4 ? __ldg(xr + d) : make_float4(0.f, 0.f, 0.f, 0.f); }
          float best = -INFINITY; int bestk = 0;
          for (int wq = 0; wq < MAX_K / 32; ++wq) {
            uint32_t mask = amb_mask[(lp * BM + r) * (MAX_K / 32) + wq];
            const int k0 = wq * 32;
            while (mask) {
              int kk[4]; int nc = 0;
#pragma unroll
              for (int c = 0; c < 4; ++c) { if (mask) { kk[c] = k0 + __ffs(mask) - 1; mask &= mask - 1; ++nc; } else kk[c] = kk[0]; }
              float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < MAXV; ++j) {
                const int d = lane + j * 32;
                if (d < D4) {
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const float4 cv = __ldg(reinterpret_cast<const float4*>(p.chat + (size_t)kk[c] * p.D) + d);
                    acc[c] = fmaf(v[j].x, cv.x, acc[c]); acc[c] = fmaf(v[j].y, cv.y, acc[c]);
                    acc[c] = fmaf(v[j].z, cv.z, acc[c]); acc[c] = fmaf(v[j].w, cv.w, acc[c]);
                  }
                }
              }
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float sc = warp_sum(acc[c]) + s_cbias[kk[c]];
                if (c < nc && sc > best) { best = sc; bestk = kk[c]; }
              }
            }
          }
          if (lane == 0) p.labels[row] = bestk;
        }
        __threadfence();
        w_sync();
        if (wt == 0) {
          // credit the tile's rows to their images; the CTA that completes an image sorts it
          const int tile = (int)blockIdx.x + arg * (int)gridDim.x;
          const int m0 = tile * BM, m1 = min(p.R, m0 + BM);
          for (int b = m0 / p.N; b <= (m1 - 1) / p.N; ++b) {
            const int c = min(m1, (b + 1) * p.N) - max(m0, b * p.N);
            if (atomicAdd(p.rows_done + b, c) + c == p.N) { sort_q[sq_tail & (SORTQ - 1)] = b; ++sq_tail; }
          }
          __threadfence();          // acquire side of the rows_done hand-over: the SORT job reads other CTAs' labels
          amb_n[lp] = 0;
          mbar_arrive(smem_u32(list_free + lp));
          ++handled;
        }
      } else if (job == J_SORT) {
        // ------------------------------------------------------------------ label sort of image `arg`, once per image
        const int b = arg, N = p.N, K = p.K;
        int* lab = arena;                                   // [N]
        float* inv = reinterpret_cast<float*>(lab + N);     // [N]
        int* cntw = reinterpret_cast<int*>(inv + N);        // [W_WARPS][K]
        int* start = cntw + W_WARPS * K;                    // [K+1]
        int* tstart = start + K + 1;                        // [K+1]
        int* sbase = tstart + K + 1;                        // [K+1]
        for (int n = wt; n < N; n += W_THREADS) {
          lab[n] = __ldcg(p.labels + (size_t)b * N + n);    // written by other CTAs in this launch: not through ld.nc
          inv[n] = p.norm_descs ? __ldcg(p.inv_norm + (size_t)b * N + n) : 1.0f;
        }
        for (int i = wt; i < W_WARPS * K; i += W_THREADS) cntw[i] = 0;
        w_sync();
        const int chunk = (((N + W_WARPS - 1) / W_WARPS) + 31) & ~31;
        const int r0 = min(N, ww * chunk), r1 = min(N, r0 + chunk);
        for (int n = r0 + lane; n < r1; n += 32) { const int l = lab[n]; if (l >= 0) atomicAdd(&cntw[ww * K + l], 1); }
        w_sync();
        for (int k = wt; k < K; k += W_THREADS) {
          int tot = 0;
          for (int q = 0; q < W_WARPS; ++q) { const int c = cntw[q * K + k]; cntw[q * K + k] = tot; tot += c; }
          start[k] = tot;
        }
        w_sync();
        if (ww == 0) {
          int run_r = 0, run_t = 0, run_s = 0;
          for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const int c = k < K ? start[k] : 0;
            const int nt = k < K ? max(1, (c + SEG - 1) / SEG) : 0;
            const int ns = nt > 1 ? nt : 0;
            int ir = c, it = nt, is = ns;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int yr = __shfl_up_sync(0xffffffffu, ir, o), yt = __shfl_up_sync(0xffffffffu, it, o), ys = __shfl_up_sync(0xffffffffu, is, o);
              if (lane >= o) { ir += yr; it += yt; is += ys; }
            }
            if (k < K) { start[k] = run_r + ir - c; tstart[k] = run_t + it - nt; sbase[k] = run_s + is - ns; }
            run_r += __shfl_sync(0xffffffffu, ir, 31); run_t += __shfl_sync(0xffffffffu, it, 31); run_s += __shfl_sync(0xffffffffu, is, 31);
          }
          if (lane == 0) { start[K] = run_r; tstart[K] = run_t; sbase[K] = run_s; }
        }
        w_sync();
        for (int k = wt; k <= K; k += W_THREADS) {
          p.t_start[(size_t)b * (K + 1) + k] = start[k];
          p.t_tstart[(size_t)b * (K + 1) + k] = tstart[k];
          p.t_sbase[(size_t)b * (K + 1) + k] = sbase[k];
        }
        for (int k = wt; k < K; k += W_THREADS)
          for (int q = tstart[k]; q < tstart[k + 1]; ++q) p.t_taskk[(size_t)b * p.max_tasks + q] = k;
        for (int n0 = r0; n0 < r1; n0 += 32) {               // stable placement, straight into the global tables
          const int n = n0 + lane;
          const int l = n < r1 ? lab[n] : -1;
          const bool active = l >= 0;
          const unsigned am = __ballot_sync(0xffffffffu, active);
          unsigned peers = 0; int rank = 0;
          if (active) {
            peers = __match_any_sync(am, l);
            rank = __popc(peers & ((1u << lane) - 1u));
            const int pos = start[l] + cntw[ww * K + l] + rank;
            p.t_ooff[(size_t)b * N + pos] = n * p.D;
            p.t_inv[(size_t)b * N + pos] = inv[n];
          }
          __syncwarp();
          if (active && rank == 0) cntw[ww * K + l] += __popc(peers);
          __syncwarp();
        }
        __threadfence();
        w_sync();
        if (wt == 0) st_release(p.image_ready + b, 1);
      } else if (job == J_ACC) {
        // ------------------------------------------------------------------ accumulate item (image, 128-column slice)
        const int b = arg / p.nslices, slice = arg - b * p.nslices, N = p.N, K = p.K, D = p.D;
        int* ooff = arena;                                  // [N]
        float* inv_s = reinterpret_cast<float*>(ooff + N);  // [N]
        int* start = reinterpret_cast<int*>(inv_s + N);     // [K+1]
        int* tstart = start + K + 1;
        int* sbase = tstart + K + 1;
        int* task_k = sbase + K + 1;                        // [max_tasks]
        float* kss = reinterpret_cast<float*>(task_k + p.max_tasks);   // [K]
        float* ksq = kss + K;                               // [K]
        float* slots = reinterpret_cast<float*>(arena) + (((size_t)2 * N + 3 * (size_t)(K + 1) + p.max_tasks + 2 * (size_t)K + 3) & ~(size_t)3);
        for (int n = wt; n < N; n += W_THREADS) {
          ooff[n] = __ldcg(p.t_ooff + (size_t)b * N + n);
          inv_s[n] = __ldcg(p.t_inv + (size_t)b * N + n);
        }
        for (int k = wt; k <= K; k += W_THREADS) {
          start[k] = __ldcg(p.t_start + (size_t)b * (K + 1) + k);
          tstart[k] = __ldcg(p.t_tstart + (size_t)b * (K + 1) + k);
          sbase[k] = __ldcg(p.t_sbase + (size_t)b * (K + 1) + k);
        }
        for (int q = wt; q < p.max_tasks; q += W_THREADS) task_k[q] = __ldcg(p.t_taskk + (size_t)b * p.max_tasks + q);
        if (wt == 0) *s_next_task = 0;
        w_sync();
        const int col = slice * 128 + lane * 4;
        const bool colok = col < D;
        const float* xb = p.x + (size_t)b * N * D + col;
        const int ntasks = tstart[K];
        auto grab = [&]() { int q = 0; if (lane == 0) q = atomicAdd(s_next_task, 1); return __shfl_sync(0xffffffffu, q, 0); };
        for (int q = grab(); q < ntasks; q = grab()) {
          const int k = task_k[q];
          const int seg = q - tstart[k], nt = tstart[k + 1] - tstart[k];
          const int s = start[k] + seg * SEG, e = min(start[k + 1], s + SEG);
          const float4 c = colok ? __ldg(reinterpret_cast<const float4*>(p.centers + (size_t)k * D + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (colok) {
            constexpr int U = 8;
            for (int i = s; i < e; i += U) {
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
            if (colok) *reinterpret_cast<float4*>(p.vlad + ((size_t)b * K + k) * D + col) = a;
            const float ss = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
            if (lane == 0) kss[k] = ss;
          } else {
            *reinterpret_cast<float4*>(slots + (size_t)(sbase[k] + seg) * 128 + lane * 4) = a;
          }
        }
        w_sync();
        for (int k = ww; k < K; k += W_WARPS) {
          const int nt = tstart[k + 1] - tstart[k];
          if (nt <= 1) continue;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < nt; ++q) {
            const float4 pv = *reinterpret_cast<const float4*>(slots + (size_t)(sbase[k] + q) * 128 + lane * 4);
            a.x += pv.x; a.y += pv.y; a.z += pv.z; a.w += pv.w;
          }
          if (colok) *reinterpret_cast<float4*>(p.vlad + ((size_t)b * K + k) * D + col) = a;
          const float ss = warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
          if (lane == 0) kss[k] = ss;
        }
        w_sync();
        for (int k = wt; k < K; k += W_THREADS) p.partial_ss[((size_t)b * K + k) * p.nslices + slice] = kss[k];
        w_sync();
        if (wt == 0) { __threadfence(); *s_last = (atomicAdd(p.done + b, 1) == p.nslices - 1); }
        w_sync();
        if (*s_last) {
          // the image's last slice: intra- and global normalisation of the whole descriptor (vlad.cu arithmetic)
          __threadfence();
          for (int k = wt; k < K; k += W_THREADS) {
            float ss = 0.f;
            for (int s = 0; s < p.nslices; ++s) ss += __ldcg(p.partial_ss + ((size_t)b * K + k) * p.nslices + s);
            const float nk = sqrtf(ss);
            const float sc = p.intra_norm ? 1.0f / fmaxf(nk, 1e-12f) : 1.0f;
            kss[k] = sc;
            const float nb = nk * sc;
            ksq[k] = nb * nb;
          }
          w_sync();
          if (wt == 0) { float tot = 0.f; for (int k = 0; k < K; ++k) tot += ksq[k]; *s_gnorm = 1.0f / fmaxf(sqrtf(tot), 1e-12f); }
          w_sync();
          const float g = *s_gnorm;
          float4* vb = reinterpret_cast<float4*>(p.vlad + (size_t)b * K * D);
          const int total4 = K * D4;
          constexpr int UN = 8;
          for (int i0 = wt; i0 < total4; i0 += W_THREADS * UN) {
            float4 v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) { const int i = i0 + u * W_THREADS; if (i < total4) v[u] = __ldcg(vb + i); }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
              const int i = i0 + u * W_THREADS;
              if (i < total4) {
                const float sc = kss[i / D4];
                v[u].x = (v[u].x * sc) * g; v[u].y = (v[u].y * sc) * g; v[u].z = (v[u].z * sc) * g; v[u].w = (v[u].w * sc) * g;
                vb[i] = v[u];
              }
            }
          }
        }
      }
      w_sync();                                               // s_job / arena are rewritten by the next round
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace v4

// Host side (draft): carve the scheduler state and the per-image tables out of `ws`, zero the state with ONE memset,
// launch one CTA per SM.  Preconditions (else the caller keeps v3): 32 <= N, K <= 128, D % 4 == 0, 32 <= D <= 2048,
// R >= 256, N * D < 2^31, 16-byte aligned feats.
size_t vlad_fused_ws_bytes(int B, int N, int K) {
  const size_t max_tasks = (size_t)N / v4::SEG + K + 1;
  return align_up((size_t)(1 + 3 * (size_t)B) * 4, 256) + 2 * align_up((size_t)B * N * 4, 256) +
         3 * align_up((size_t)B * (K + 1) * 4, 256) + align_up((size_t)B * max_tasks * 4, 256) + 1024;
}

int vlad_fused_launch(const float* feats, const int32_t* n_valid, const float* centers, const float* chat,
                      const float* chat_tf32, const float* cbias, const float* cnorm, const float* cdnorm, int B, int N,
                      int D, int K, int norm_descs, int intra_norm, int32_t* labels, float* inv_norm, float* vlad,
                      float* partial_ss, void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace v4;
  const int64_t R64 = (int64_t)B * N;
  ANYLOC_REQUIRE(N >= 32 && K >= 1 && K <= MAX_K && D >= 32 && D <= 2048 && D % 4 == 0 && R64 >= 256 && R64 < (1ll << 31) - 256 &&
                 (int64_t)N * D < (1ll << 31), "vlad_fused: unsupported shape B=%d N=%d D=%d K=%d", B, N, D, K);
  Workspace w(ws, ws_bytes);
  Params p;
  int32_t* state = w.take<int32_t>(1 + 3 * (size_t)B);
  p.t_ooff = w.take<int32_t>((size_t)B * N);
  p.t_inv = w.take<float>((size_t)B * N);
  p.t_start = w.take<int32_t>((size_t)B * (K + 1));
  p.t_tstart = w.take<int32_t>((size_t)B * (K + 1));
  p.t_sbase = w.take<int32_t>((size_t)B * (K + 1));
  p.max_tasks = N / SEG + K + 1;
  p.max_slots = 2 * (N / SEG) + 2;
  p.t_taskk = w.take<int32_t>((size_t)B * p.max_tasks);
  if (!state || !p.t_taskk) { set_error("vlad_fused: workspace too small"); return ANYLOC_ERR_WORKSPACE; }
  p.next_acc = state; p.rows_done = state + 1; p.image_ready = p.rows_done + B; p.done = p.image_ready + B;
  ANYLOC_CHECK_CUDA(cudaMemsetAsync(state, 0, (1 + 3 * (size_t)B) * 4, st));
  p.x = feats; p.n_valid = n_valid; p.B = B; p.N = N; p.D = D; p.K = K; p.R = (int)R64; p.centers = centers;
  p.chat = chat; p.cbias = cbias; p.cnorm = cnorm; p.cdnorm = cdnorm; p.norm_descs = norm_descs; p.intra_norm = intra_norm;
  p.labels = labels; p.inv_norm = inv_norm; p.vlad = vlad; p.partial_ss = partial_ss;
  p.n_mma = (K + 15) / 16 * 16;
  p.stage_bytes = A_BYTES + p.n_mma * 128;
  p.nslices = cdiv(D, 128);
  p.num_tiles = (int)((R64 + BM - 1) / BM);
  // W-side arena: the larger of the SORT and the ACC working sets
  const size_t sort_ints = (size_t)2 * N + (size_t)W_WARPS * K + 3 * (size_t)(K + 1);
  const size_t acc_ints = (size_t)2 * N + 3 * (size_t)(K + 1) + p.max_tasks + 2 * (size_t)K + 4;
  const size_t arena = std::max(sort_ints * 4, acc_ints * 4 + (size_t)p.max_slots * 512) + 64;
  int dev = 0, max_smem = 0;
  ANYLOC_CHECK_CUDA(cudaGetDevice(&dev));
  ANYLOC_CHECK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = 1024 + BAR_BYTES + VEC_BYTES + LIST_BYTES + 64 + arena;
  ANYLOC_REQUIRE(fixed + 2 * (size_t)p.stage_bytes <= (size_t)max_smem, "vlad_fused: N=%d K=%d need %zu B of shared memory", N, K, fixed);
  p.stages = (int)std::min<size_t>(MAX_STAGES, ((size_t)max_smem - fixed) / p.stage_bytes);
  const size_t smem = (size_t)p.stages * p.stage_bytes + fixed;
  CUtensorMap mx, mc;
  int rc;
  if ((rc = tc::make_map(&mx, feats, p.R, D, D, BM, false))) return rc;
  if ((rc = tc::make_map(&mc, chat_tf32, K, D, D, p.n_mma, false))) return rc;
  const int grid = device_sm_count();                         // CTAs without tiles still work as accumulators
#define ANYLOC_LAUNCH_V4(MAXV_)                                                                                          \
  do {                                                                                                                    \
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(vlad_fused_kernel<MAXV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)); \
    vlad_fused_kernel<MAXV_><<<grid, THREADS, smem, st>>>(mx, mc, p);                                                     \
  } while (0)
  if (D <= 512) ANYLOC_LAUNCH_V4(4); else if (D <= 1024) ANYLOC_LAUNCH_V4(8); else ANYLOC_LAUNCH_V4(16);
#undef ANYLOC_LAUNCH_V4
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

}  // namespace anyloc
