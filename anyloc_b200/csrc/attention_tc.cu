// tcgen05 flash attention for the DINOv2 blocks (upstream dinov2/layers/attention.py via
// /root/reference/utilities.py:269): O = softmax(Q K^T / 8) V per head (head_dim 64), fp32-equivalent
// accuracy through the 3-term tf32 split on both GEMMs.
//
// Inputs are the (hi,lo) tf32 pairs of the fused qkv projection: q,k thirds in [B*T, 3D], and V stored
// per-head TRANSPOSED, vt[(b*heads+h)*64 + d][t] (row pitch Tp), both written by the qkv GEMM's QKV_SPLIT
// epilogue.  (tcgen05 MN-major B operands need the 32B-atom swizzle for 32-bit types -- measured: an
// MN-major tf32 operand in the plain 128B swizzle yields zeros -- so V is made K-major at the source.)
// Output is the (hi,lo) pair [B*T, D] that feeds the `proj` GEMM.
// One CTA per (128-query tile, head, image), 256 threads:
//   (persistent: CTAs stride over the (q-tile, head, image) work items so TMEM allocation, barrier init and
//    load latency are paid once and the producer/MMA warps run ahead into the next item)
//   warp 0/3 : TMA producers -- Q tile per item + 2-stage K ring (warp 0), 2-stage V^T ring (warp 3), 64-key blocks
//   warp 1   : MMA issuer   -- S_j = Q K_j^T  (A,B from smem, K-major, M128 x N64 x K8, 24 UMMAs)
//                              O_j = P_j V_j  (A = P from TMEM, B = V^T tile from smem K-major, 24 UMMAs)
//   warp 2   : TMEM allocator (S double-buffered 2x64, P_hi 64, P_lo 64, O chunk 64 columns)
//   warps 4-11: softmax     -- two warps per 32-row TMEM lane quarter (key / head-dim halves): tcgen05.ld S,
//                              online softmax in fp32 registers,
//                              tcgen05.st P (hi,lo), and round-to-nearest accumulation of the per-block
//                              O_j chunks into register accumulators (the tensor core's accumulator
//                              rounds toward zero, see gemm_tc.cu), final 1/l scaling, (hi,lo) stores.
#include <cuda.h>
#include "common.cuh"

namespace anyloc {
namespace atc {

constexpr int BQ = 128;        // queries per CTA
constexpr int BKV = 64;        // keys per block
constexpr int HD = 64;         // head dim
constexpr int STAGES = 2;
constexpr int Q_HALF = BQ * 32 * 4;           // one [128 x 32] fp32 k-block: 16 KB
constexpr int Q_BYTES = 4 * Q_HALF;           // hi(2 k-blocks) + lo(2 k-blocks): 64 KB
constexpr int KV_BOX = BKV * 32 * 4;          // one [64 x 32] fp32 box: 8 KB
constexpr int STAGE_BYTES = 8 * KV_BOX;       // K_hi(2) K_lo(2) V_hi(2) V_lo(2): 64 KB
constexpr int THREADS = 128 + 256;        // producer / MMA / alloc / spare + 8 softmax warps
constexpr int XCHG_BYTES = 6 * BQ * 4;    // row-max (2 slots x 2 halves) and row-sum (2 halves) exchange
constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + 1024 + 256 + XCHG_BYTES;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0;       // 2 x 64
constexpr uint32_t COL_PHI = 128;   // 64
constexpr uint32_t COL_PLO = 192;   // 64
constexpr uint32_t COL_O = 256;     // 64

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && (it & 0x3ff) == 0x3ff) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand, 128B swizzle (8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_hi_q, const __grid_constant__ CUtensorMap tm_lo_q,
                    const __grid_constant__ CUtensorMap tm_hi_kv, const __grid_constant__ CUtensorMap tm_lo_kv,
                    const __grid_constant__ CUtensorMap tm_hi_vt, const __grid_constant__ CUtensorMap tm_lo_vt,
                    int B, int T, int D, float* __restrict__ o_hi, float* __restrict__ o_lo, int out_f16,
                    float* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                  // [hi kb0][hi kb1][lo kb0][lo kb1], 16 KB each
  uint8_t* sKV = smem + Q_BYTES;                       // STAGES x {K_hi0,K_hi1,K_lo0,K_lo1,Vt_hi0,Vt_hi1,Vt_lo0,Vt_lo1}
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * STAGE_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* q_empty = bars + 1;            // 1
  uint64_t* k_full = bars + 2;             // [STAGES]   K and V^T rings are released separately: K_j as soon
  uint64_t* k_empty = k_full + STAGES;     // [STAGES]   as S_j retires (early), V_j after PV_j -- so the next
  uint64_t* v_full = k_empty + STAGES;     // [STAGES]   K block streams in while softmax/PV of the current
  uint64_t* v_empty = v_full + STAGES;     // [STAGES]   block are still running
  uint64_t* s_full = v_empty + STAGES;     // [2]
  uint64_t* p_full = s_full + 2;           // 1
  uint64_t* o_full = p_full + 1;           // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xchg = reinterpret_cast<float*>(bars) + 64;      // 256 B after the barrier block

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int heads = D / HD;
  const int q_tiles = (T + BQ - 1) / BQ;
  const int nblk = (T + BKV - 1) / BKV;
  const int total = q_tiles * heads * B;       // work items; persistent CTAs stride over them

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi_vt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo_vt) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(smem_u32(q_full), 1); mbar_init(smem_u32(q_empty), 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(k_full + s), 1); mbar_init(smem_u32(k_empty + s), 1);
      mbar_init(smem_u32(v_full + s), 1); mbar_init(smem_u32(v_empty + s), 1);
    }
    mbar_init(smem_u32(s_full), 1); mbar_init(smem_u32(s_full + 1), 1);
    mbar_init(smem_u32(p_full), 8);          // one arrive per softmax warp
    mbar_init(smem_u32(o_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // All three roles walk the same sequence of work items and of key blocks; g counts key blocks globally so the
  // ring / double-buffer parities carry over from one work item to the next.
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer: Q tiles and the K ring
      int g = 0, it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int qt = w % q_tiles, h = (w / q_tiles) % heads, b = w / (q_tiles * heads);
        const int row0 = b * T, colq = h * HD, colk = D + h * HD;
        mbar_wait(smem_u32(q_empty), (uint32_t)((it & 1) ^ 1));     // previous tile's S MMAs are done with Q
        const uint32_t qb = smem_u32(q_full);
        mbar_expect_tx(qb, Q_BYTES);
        tma_load_2d(smem_u32(sQ + 0 * Q_HALF), &tm_hi_q, qb, colq, row0 + qt * BQ);
        tma_load_2d(smem_u32(sQ + 1 * Q_HALF), &tm_hi_q, qb, colq + 32, row0 + qt * BQ);
        tma_load_2d(smem_u32(sQ + 2 * Q_HALF), &tm_lo_q, qb, colq, row0 + qt * BQ);
        tma_load_2d(smem_u32(sQ + 3 * Q_HALF), &tm_lo_q, qb, colq + 32, row0 + qt * BQ);
        for (int j = 0; j < nblk; ++j, ++g) {
          const int stage = g % STAGES;
          mbar_wait(smem_u32(k_empty + stage), (uint32_t)(((g / STAGES) & 1) ^ 1));
          const uint32_t fb = smem_u32(k_full + stage);
          mbar_expect_tx(fb, STAGE_BYTES / 2);
          const uint32_t sb = smem_u32(sKV + stage * STAGE_BYTES);
          const int r = row0 + j * BKV;
          tma_load_2d(sb + 0 * KV_BOX, &tm_hi_kv, fb, colk, r);
          tma_load_2d(sb + 1 * KV_BOX, &tm_hi_kv, fb, colk + 32, r);
          tma_load_2d(sb + 2 * KV_BOX, &tm_lo_kv, fb, colk, r);
          tma_load_2d(sb + 3 * KV_BOX, &tm_lo_kv, fb, colk + 32, r);
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer: the V^T ring
      int g = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int h = (w / q_tiles) % heads, b = w / (q_tiles * heads);
        const int vrow = (b * heads + h) * HD;
        for (int j = 0; j < nblk; ++j, ++g) {
          const int stage = g % STAGES;
          mbar_wait(smem_u32(v_empty + stage), (uint32_t)(((g / STAGES) & 1) ^ 1));
          const uint32_t fb = smem_u32(v_full + stage);
          mbar_expect_tx(fb, STAGE_BYTES / 2);
          const uint32_t sb = smem_u32(sKV + stage * STAGE_BYTES);
          tma_load_2d(sb + 4 * KV_BOX, &tm_hi_vt, fb, j * BKV, vrow);         // V^T [64 d x 32 keys] k-block 0
          tma_load_2d(sb + 5 * KV_BOX, &tm_hi_vt, fb, j * BKV + 32, vrow);    //                      k-block 1
          tma_load_2d(sb + 6 * KV_BOX, &tm_lo_vt, fb, j * BKV, vrow);
          tma_load_2d(sb + 7 * KV_BOX, &tm_lo_vt, fb, j * BKV + 32, vrow);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer
      // S: M128 x N64 (keys), A/B K-major.  PV: M128 x N64 (head dim), A = P from TMEM, B = V^T K-major.
      constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BKV >> 3) << 17) |
                                   ((uint32_t)(BQ >> 4) << 24);
      constexpr uint32_t idesc_pv = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HD >> 3) << 17) |
                                    ((uint32_t)(BQ >> 4) << 24);
      const uint32_t q_base = smem_u32(sQ);
      auto issue_s = [&](int gb) {
        const int st = gb % STAGES;
        mbar_wait(smem_u32(k_full + st), (uint32_t)((gb / STAGES) & 1));
        tc_fence_after();
        const uint32_t kb = smem_u32(sKV + st * STAGE_BYTES);
        const uint32_t d = tmem_base + COL_S + (uint32_t)((gb & 1) * BKV);
#pragma unroll
        for (int k = 0; k < HD / 8; ++k) {                   // 8 k-steps over the head dim
          const uint32_t off = (uint32_t)((k >> 2) * Q_HALF + (k & 3) * 32);
          const uint32_t koff = (uint32_t)((k >> 2) * KV_BOX + (k & 3) * 32);
          const uint64_t a_hi = desc_kmajor(q_base + off), a_lo = desc_kmajor(q_base + 2 * Q_HALF + off);
          const uint64_t b_hi = desc_kmajor(kb + koff), b_lo = desc_kmajor(kb + 2 * KV_BOX + koff);
          umma_ss(d, a_hi, b_hi, idesc_s, k != 0);
          umma_ss(d, a_lo, b_hi, idesc_s, 1u);
          umma_ss(d, a_hi, b_lo, idesc_s, 1u);
        }
        umma_commit(smem_u32(s_full + (gb & 1)));
        umma_commit(smem_u32(k_empty + st));      // K_gb may be overwritten as soon as S_gb retires
      };
      int g = 0, it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const bool tdump = dbg != nullptr && it == 1 && blockIdx.x == 0;
        const long long tb = tdump ? clock64() : 0;
#define MSTAMP(jj, slot) do { if (tdump) dbg[40960 + (jj) * 16 + 8 + (slot)] = (float)(clock64() - tb); } while (0)
        mbar_wait(smem_u32(q_full), (uint32_t)(it & 1));
        tc_fence_after();
        MSTAMP(0, 0);
        issue_s(g);
        MSTAMP(0, 1);
        if (nblk == 1) umma_commit(smem_u32(q_empty));
        for (int j = 0; j < nblk; ++j) {
          const int gb = g + j, st = gb % STAGES;
          if (j + 1 < nblk) {
            MSTAMP(j + 1, 0);
            issue_s(gb + 1);       // S buffer (gb+1)&1 was consumed before P_{gb-1} was published (program order)
            MSTAMP(j + 1, 1);
            if (j + 2 == nblk) umma_commit(smem_u32(q_empty));   // last S of this tile issued: Q may be reloaded
          }
          mbar_wait(smem_u32(v_full + st), (uint32_t)((gb / STAGES) & 1));
          mbar_wait(smem_u32(p_full), (uint32_t)(gb & 1));
          tc_fence_after();
          MSTAMP(j, 2);
          const uint32_t vb = smem_u32(sKV + st * STAGE_BYTES) + 4 * KV_BOX;
          const uint32_t d = tmem_base + COL_O;
#pragma unroll
          for (int k = 0; k < BKV / 8; ++k) {                   // 8 k-steps over the 64 keys
            const uint32_t voff = (uint32_t)((k >> 2) * KV_BOX + (k & 3) * 32);
            const uint64_t v_hi = desc_kmajor(vb + voff), v_lo = desc_kmajor(vb + 2 * KV_BOX + voff);
            const uint32_t p_hi = tmem_base + COL_PHI + (uint32_t)(k * 8), p_lo = tmem_base + COL_PLO + (uint32_t)(k * 8);
            umma_ts(d, p_hi, v_hi, idesc_pv, k != 0);
            umma_ts(d, p_lo, v_hi, idesc_pv, 1u);
            umma_ts(d, p_hi, v_lo, idesc_pv, 1u);
          }
          umma_commit(smem_u32(o_full));
          MSTAMP(j, 3);
          umma_commit(smem_u32(v_empty + st));      // V_gb no longer needed once these MMAs retire
        }
        g += nblk;
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------- softmax + RN accumulation
    // 8 warps: two per TMEM lane quarter.  Both own the same 32 query rows; warp `half` handles key columns
    // [32*half, 32*half+32) of every S block and head-dim columns [32*half, +32) of O, so each thread carries
    // 32 + 32 live values instead of 64 + 64 and every SM sub-partition has two warps to hide latency with.
    // Only the running row maximum must agree between the pair: exchanged through shared memory per block.
    const int qd = warp & 3, half = (warp - 4) >> 2;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
    int g = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int qt = w % q_tiles, h = (w / q_tiles) % heads, b = w / (q_tiles * heads);
      const int qrow = qt * BQ + row;                     // token index inside the image
      const bool dump = dbg != nullptr && w == 0;
      const bool tdump = dbg != nullptr && w == (int)gridDim.x && blockIdx.x == 0 && warp == 4 && lane == 0;  // 2nd item of CTA 0
      long long tb = tdump ? clock64() : 0;
#define TSTAMP(slot) do { if (tdump) dbg[40960 + j * 16 + (slot)] = (float)(clock64() - tb); } while (0)
      float m = -INFINITY, l = 0.f;                       // l: this warp's half of the row sum
      float o[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const int gb = g + j;
        TSTAMP(0);
        mbar_wait(smem_u32(s_full + (gb & 1)), (uint32_t)((gb >> 1) & 1));
        tc_fence_after();
        TSTAMP(1);
        float s[32];
        tmem_ld32(lane_addr + COL_S + (uint32_t)((gb & 1) * BKV + half * 32), s);
        TSTAMP(2);
        if (j == nblk - 1) {                              // only the last block can hold keys >= T
#pragma unroll
          for (int c = 0; c < 32; ++c) if (j * BKV + half * 32 + c >= T) s[c] = -INFINITY;
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          s[c] *= kScale; s[c + 1] *= kScale; s[c + 2] *= kScale; s[c + 3] *= kScale;
          mx0 = fmaxf(mx0, s[c]); mx1 = fmaxf(mx1, s[c + 1]); mx2 = fmaxf(mx2, s[c + 2]); mx3 = fmaxf(mx3, s[c + 3]);
        }
        const float pm = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        xchg[((gb & 1) * 2 + half) * BQ + row] = pm;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");      // the two warps of this lane quarter
        const float mx = fmaxf(m, fmaxf(pm, xchg[((gb & 1) * 2 + (half ^ 1)) * BQ + row]));
        TSTAMP(3);
        if (dump && j == 0) for (int c = 0; c < 32; ++c) dbg[row * 64 + half * 32 + c] = s[c];
        const float alpha = ex2(m - mx);                  // 0 on the first block (m = -inf, mx finite)
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          s[c] = ex2(s[c] - mx); s[c + 1] = ex2(s[c + 1] - mx);
          s[c + 2] = ex2(s[c + 2] - mx); s[c + 3] = ex2(s[c + 3] - mx);
          r0 += s[c]; r1 += s[c + 1]; r2 += s[c + 2]; r3 += s[c + 3];
        }
        l = l * alpha + ((r0 + r1) + (r2 + r3));
        m = mx;
        TSTAMP(4);
        if (dump && j == 0) for (int c = 0; c < 32; ++c) dbg[8192 + row * 64 + half * 32 + c] = s[c];
        if (j > 0) {                                      // fold in O_{j-1} (RN), frees the P and O buffers
          mbar_wait(smem_u32(o_full), (uint32_t)((gb - 1) & 1));
          tc_fence_after();
          TSTAMP(5);
          float t[32];
          tmem_ld32(lane_addr + COL_O + (uint32_t)(half * 32), t);
          if (dump && j == 1) for (int c = 0; c < 32; ++c) dbg[16384 + row * 64 + half * 32 + c] = t[c];
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = (o[c] + t[c]) * alpha;
        }
        // publish this warp's 32 key columns of P_j = (hi, lo)
        {
          float t[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) { float hh, ll; split_tf32(s[c], hh, ll); t[c] = hh; s[c] = ll; }
          tmem_st32(lane_addr + COL_PHI + (uint32_t)(half * 32), t);
          tmem_st32(lane_addr + COL_PLO + (uint32_t)(half * 32), s);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        TSTAMP(6);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(p_full));
        TSTAMP(7);
      }
      // last chunk of this work item, then the row sum of the partner warp
      mbar_wait(smem_u32(o_full), (uint32_t)((g + nblk - 1) & 1));
      tc_fence_after();
      {
        float t[32];
        tmem_ld32(lane_addr + COL_O + (uint32_t)(half * 32), t);
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] += t[c];
      }
      g += nblk;
      xchg[(4 + half) * BQ + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
      const float inv = 1.0f / (l + xchg[(4 + (half ^ 1)) * BQ + row]);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");        // slot reusable by the next work item
      if (qrow < T) {
        const size_t off = ((size_t)b * T + qrow) * D + (size_t)h * HD + half * 32;
        if (out_f16) {
          uint4* ph = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(o_hi) + off);
          uint4* pl = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(o_lo) + off);
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            __half hh[8], ll[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) split_f16(o[c + q] * inv * kActScale, hh[q], ll[q]);
            ph[c >> 3] = *reinterpret_cast<uint4*>(hh); pl[c >> 3] = *reinterpret_cast<uint4*>(ll);
          }
        } else {
          float4* ph = reinterpret_cast<float4*>(o_hi + off);
          float4* pl = reinterpret_cast<float4*>(o_lo + off);
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            float4 hh, ll;
            split_tf32(o[c] * inv, hh.x, ll.x); split_tf32(o[c + 1] * inv, hh.y, ll.y);
            split_tf32(o[c + 2] * inv, hh.z, ll.z); split_tf32(o[c + 3] * inv, hh.w, ll.w);
            ph[c >> 2] = hh; pl[c >> 2] = ll;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
static int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("attention_tc: cuTensorMapEncodeTiled unavailable"); return ANYLOC_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return ANYLOC_ERR_CUDA; }
  return ANYLOC_OK;
}

// V third of qkv -> per-head transposed (hi,lo): vt[(b*heads+h)*64 + d][t], row pitch Tp.  (Standalone
// building-block path only; inside the ViT the qkv GEMM epilogue writes this layout directly.)
__global__ void __launch_bounds__(128)
v_transpose_kernel(const float* __restrict__ qkv_hi, const float* __restrict__ qkv_lo, int T, int Tp, int D,
                   float* __restrict__ vt_hi, float* __restrict__ vt_lo) {
  const int t = blockIdx.x * 128 + threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const size_t src = ((size_t)b * T + t) * 3 * D + 2 * D + (size_t)h * HD;
  const size_t dst = ((size_t)b * D + (size_t)h * HD) * Tp + t;
#pragma unroll 4
  for (int d = 0; d < HD; d += 4) {
    float4 a = __ldg(reinterpret_cast<const float4*>(qkv_hi + src + d));
    float4 c = __ldg(reinterpret_cast<const float4*>(qkv_lo + src + d));
    vt_hi[dst + (size_t)(d + 0) * Tp] = a.x; vt_hi[dst + (size_t)(d + 1) * Tp] = a.y;
    vt_hi[dst + (size_t)(d + 2) * Tp] = a.z; vt_hi[dst + (size_t)(d + 3) * Tp] = a.w;
    vt_lo[dst + (size_t)(d + 0) * Tp] = c.x; vt_lo[dst + (size_t)(d + 1) * Tp] = c.y;
    vt_lo[dst + (size_t)(d + 2) * Tp] = c.z; vt_lo[dst + (size_t)(d + 3) * Tp] = c.w;
  }
}

}  // namespace atc

int attention_vt_pitch(int T) { return (T + 3) & ~3; }     // TMA row pitch must be a multiple of 16 bytes

// vt_{hi,lo}: [B*D, Tp] with columns [T, Tp) zero (cudaMemset once; never written afterwards)
int attention_tc_launch(const float* qkv_hi, const float* qkv_lo, const float* vt_hi, const float* vt_lo, int B, int T,
                        int D, int heads, void* o_hi, void* o_lo, bool out_f16, float* dbg, cudaStream_t st) {
  using namespace atc;
  ANYLOC_REQUIRE(D == heads * HD, "attention_tc: head_dim must be 64 (D=%d heads=%d)", D, heads);
  CUtensorMap hq, lq, hkv, lkv, hvt, lvt;
  int rc;
  const int64_t rows = (int64_t)B * T;
  const int Tp = attention_vt_pitch(T);
  if ((rc = make_map(&hq, qkv_hi, rows, 3 * D, BQ))) return rc;
  if ((rc = make_map(&lq, qkv_lo, rows, 3 * D, BQ))) return rc;
  if ((rc = make_map(&hkv, qkv_hi, rows, 3 * D, BKV))) return rc;
  if ((rc = make_map(&lkv, qkv_lo, rows, 3 * D, BKV))) return rc;
  if ((rc = make_map(&hvt, vt_hi, (int64_t)B * D, Tp, HD))) return rc;
  if ((rc = make_map(&lvt, vt_lo, (int64_t)B * D, Tp, HD))) return rc;
  static unsigned long long attr_seen = 0;
  if (first_use_on_this_device(&attr_seen)) {
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  const int total = cdiv(T, BQ) * heads * B;
  attention_tc_kernel<<<std::min(total, device_sm_count()), THREADS, SMEM_BYTES, st>>>(
      hq, lq, hkv, lkv, hvt, lvt, B, T, D, (float*)o_hi, (float*)o_lo, out_f16 ? 1 : 0, dbg);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

// standalone: transposes V into a stream-ordered temporary first
int attention_tc_standalone(const float* qkv_hi, const float* qkv_lo, int B, int T, int D, int heads, void* o_hi,
                            void* o_lo, bool out_f16, float* dbg, cudaStream_t st) {
  const int Tp = attention_vt_pitch(T);
  const size_t n = (size_t)B * D * Tp;
  float* vt = nullptr;
  ANYLOC_CHECK_CUDA(cudaMallocAsync((void**)&vt, 2 * n * sizeof(float), st));
  ANYLOC_CHECK_CUDA(cudaMemsetAsync(vt, 0, 2 * n * sizeof(float), st));
  atc::v_transpose_kernel<<<dim3(cdiv(T, 128), heads, B), 128, 0, st>>>(qkv_hi, qkv_lo, T, Tp, D, vt, vt + n);
  ANYLOC_CHECK_LAUNCH();
  int rc = attention_tc_launch(qkv_hi, qkv_lo, vt, vt + n, B, T, D, heads, o_hi, o_lo, out_f16, dbg, st);
  cudaFreeAsync(vt, st);
  return rc;
}

}  // namespace anyloc

// debug entry (not part of the public ABI): dumps S/P/O-chunk of block 0 of CTA (0,0,0) into dbg[3*8192]
extern "C" int anyloc_attention_tc_debug(const float* qkv_hi, const float* qkv_lo, int B, int T, int D, int heads,
                                         float* o_hi, float* o_lo, float* dbg, void* stream) {
  return anyloc::attention_tc_standalone(qkv_hi, qkv_lo, B, T, D, heads, o_hi, o_lo, false, dbg, (cudaStream_t)stream);
}
