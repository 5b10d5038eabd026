// tcgen05 GEMM engine:  C[M,N] = (A_hi+A_lo)[M,K] . (B_hi+B_lo)[N,K]^T  + fused epilogue,
// fp32-equivalent accuracy through the 3-term tf32 split
//      A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo        (all accumulated in fp32 in TMEM)
// where x_hi = rna_tf32(x), x_lo = x - x_hi are materialised by the producer kernels
// (LayerNorm / previous epilogue / weight prep), so the tensor core consumes plain fp32 words.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D, 128B-swizzled K-major boxes of 32 floats
//   warp 1   : MMA issuer    -- one lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x BN x K8)
//   warp 2   : TMEM allocator (2 x BN fp32 columns: double-buffered CHUNK accumulator)
//   warps 4-19: accumulate + epilogue -- the tensor core adds into its fp32 accumulator with
//               round-toward-zero, a bias of ~2^-26 per MMA that grows linearly with K (measured:
//               -9e-6 relative at K=1536, -1e-4 at K=16384).  So the MMA warp only accumulates
//               CHUNK_KB k-blocks (K=64) in TMEM; these warps drain every chunk with tcgen05.ld and
//               add it to round-to-nearest fp32 register accumulators (the whole 128xBN tile lives
//               in registers: 64 columns of one row per thread), which brings the error back to
//               the level of an fp32 FFMA GEMM.  After the last chunk: bias / GELU / SwiGLU /
//               LayerScale+residual / tf32 split and vectorised global stores, overlapping the next
//               tile's MMAs.
// Tiles are rastered n-fastest so that the CTAs resident at any moment share a handful of A row
// panels and the whole B matrix in L2.
#include <cuda.h>
#include <stdlib.h>
#include "epilogue.cuh"
#include "tc_common.cuh"

namespace anyloc {

namespace tc {

constexpr int BM = 128;
constexpr int KSTEPS = 4;                  // UMMA k-steps per 128-byte k-block (32 B each: 8 tf32 or 16 fp16)
constexpr int A_BYTES = BM * 128;         // 16 KB: 128 rows x 128 B

// LO = true : stages hold {A_hi, A_lo, B_hi, B_lo} (3-term split).  LO = false: hi-only single pass (coarse
// scores, e.g. the VLAD assignment): {A_hi, B_hi} per stage -> twice the pipeline depth for latency-bound streams.
template <int BN, bool LO = true> struct Cfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = (LO ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int B_OFF = (LO ? 2 : 1) * A_BYTES;
  static constexpr int STAGES = (LO ? 2 : 4);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;          // 512 or 256: power of two
};
constexpr int CHUNK_KB_TF32 = 2;          // k-blocks (128 B of K each) accumulated in the tensor core before an RN
constexpr int CHUNK_KB_F16 = 4;           // drain: 24 / 48 MMAs per chunk -> RZ bias ~4e-7 / ~8e-7 relative (1-CTA kernel)
constexpr int CHUNK_KB_F16_2CTA = 8;      // the ViT's 2-CTA fp16 kernel: 96 MMAs per chunk -- measured at c2 (31 blocks): features 1.1e-5
                                          // from the oracle instead of 6.1e-6 (tolerance 1e-4) for +1.4 % throughput: half the TMEM drains
constexpr int EPI_WARPS = 16;             // 4 TMEM lane quarters x 4 column quarters
constexpr int THREADS = 128 + EPI_WARPS * 32;

// Tile raster: bands of BAND_N column blocks; inside a band the tiles run n-fastest over all row blocks.  The ~148
// resident CTAs then share a few A row panels and ONE band of B (<= BAND_N*256 rows of hi+lo) that stays in L2 while
// the outputs stream through it; with a plain n-fastest raster over a wide N the whole B matrix (50-100 MB) is
// evicted and re-read from HBM by every wave (ncu: up to 10x the algorithmic DRAM traffic on the w12 GEMM).
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int band_n, int& m_blk, int& n_blk) {
  const int per_band = num_m * band_n;
  const int band = tile / per_band, r = tile - band * per_band;
  const int w = min(band_n, num_n - band * band_n);       // width of this (possibly last, narrower) band
  m_blk = r / w;
  n_blk = band * band_n + (r - m_blk * w);
}


// ---- epilogue on 32 consecutive columns of one row (row m, columns n..n+31), raw accumulators in v[]
__device__ __forceinline__ void store_split4(const EpiParams& ep, size_t o, float4 x) {
  if (ep.out_f16) {      // fp16 pair of kActScale*x: 4 halves = 8 bytes per array
    uint2 h, l;
    split_f16x2(x.x * kActScale, x.y * kActScale, h.x, l.x);
    split_f16x2(x.z * kActScale, x.w * kActScale, h.y, l.y);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out) + o) = h;
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out_lo) + o) = l;
  } else {
    float4 h, l;
    split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
    *reinterpret_cast<float4*>(ep.out + o) = h;
    *reinterpret_cast<float4*>(ep.out_lo + o) = l;
  }
}

// MODE: the epilogue mode as a compile-time constant (the 2-CTA kernel is instantiated per mode so that the code a
// tile executes stays compact), or -2 = read ep.mode at run time.
template <int MODE = -2>
__device__ __forceinline__ void epi_chunk32(const EpiParams& ep_in, int m, int n, int N, const float* v) {
  const int mode = MODE == -2 ? ep_in.mode : MODE;
  if (mode < 0) return;                    // diagnostic: discard (ANYLOC_GEMM_DEBUG_SKIP_EPI)
  const EpiParams& ep = ep_in;
  const bool vec = (n + 32 <= N) && ((ep.ldo & 3) == 0);
  if (!vec) {
    if (mode == ANYLOC_EPI_SWIGLU_SPLIT) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) if (n + j + 1 < N) epi_store_pair(ep, m, n + j, v[j], v[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (n + j < N) epi_store1(ep, m, n + j, v[j]);
    }
    return;
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float al = ep.alpha;
  if (mode == ANYLOC_EPI_QKV_SPLIT && n >= 2 * ep.qkv_D) {
    // V third: transposed per-head store; lanes of a warp hold consecutive rows -> coalesced along t
    const int c0 = n - 2 * ep.qkv_D, b = m / ep.qkv_T, t = m - b * ep.qkv_T;
    const size_t o0 = ((size_t)b * ep.qkv_D + c0) * ep.qkv_Tp + t;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = v[j] * al + (ep.bias ? __ldg(ep.bias + n + j) : 0.f);
      if (ep.qkv_f16) {
        __half h, l; split_f16(x * kActScale, h, l);
        reinterpret_cast<__half*>(ep.vt_hi)[o0 + (size_t)j * ep.qkv_Tp] = h;
        reinterpret_cast<__half*>(ep.vt_lo)[o0 + (size_t)j * ep.qkv_Tp] = l;
      } else {
        float h, l; split_tf32(x, h, l);
        ep.vt_hi[o0 + (size_t)j * ep.qkv_Tp] = h;
        ep.vt_lo[o0 + (size_t)j * ep.qkv_Tp] = l;
      }
    }
    return;
  }
  if (mode == ANYLOC_EPI_SWIGLU_SPLIT) {
    const size_t o = (size_t)m * ep.ldo + (n >> 1);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float4 b0 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n + j)) : zero4;
      float4 b1 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n + j + 4)) : zero4;
      float4 x;
      x.x = silu_fast(v[j] * al + b0.x) * (v[j + 1] * al + b0.y);
      x.y = silu_fast(v[j + 2] * al + b0.z) * (v[j + 3] * al + b0.w);
      x.z = silu_fast(v[j + 4] * al + b1.x) * (v[j + 5] * al + b1.y);
      x.w = silu_fast(v[j + 6] * al + b1.z) * (v[j + 7] * al + b1.w);
      store_split4(ep, o + (j >> 1), x);
    }
    return;
  }
  const size_t o = (size_t)m * ep.ldo + n;
  const bool qkv = mode == ANYLOC_EPI_QKV_SPLIT;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float4 b = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n + j)) : zero4;
    float4 x = make_float4(v[j] * al + b.x, v[j + 1] * al + b.y, v[j + 2] * al + b.z, v[j + 3] * al + b.w);
    if (mode == ANYLOC_EPI_BIAS) {
      reinterpret_cast<float4*>(ep.out + o)[j >> 2] = x;
    } else if (mode == ANYLOC_EPI_LS_RESID) {
      float4 r = reinterpret_cast<const float4*>(ep.resid + o)[j >> 2];
      float4 g = __ldg(reinterpret_cast<const float4*>(ep.gamma + n + j));
      reinterpret_cast<float4*>(ep.out + o)[j >> 2] =
          make_float4(r.x + g.x * x.x, r.y + g.y * x.y, r.z + g.z * x.z, r.w + g.w * x.w);
    } else if (qkv && ep.qkv_f16) {   // q,k thirds as fp16 pairs (f16 attention input)
      uint2 h, l;
      split_f16x2(x.x * kActScale, x.y * kActScale, h.x, l.x);
      split_f16x2(x.z * kActScale, x.w * kActScale, h.y, l.y);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out) + o + j) = h;
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out_lo) + o + j) = l;
    } else if (qkv) {   // q,k thirds as tf32 pairs (tf32 attention input)
      float4 h, l;
      split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
      reinterpret_cast<float4*>(ep.out + o)[j >> 2] = h;
      reinterpret_cast<float4*>(ep.out_lo + o)[j >> 2] = l;
    } else {            // BIAS_SPLIT / GELU_SPLIT
      if (mode == ANYLOC_EPI_GELU_SPLIT) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
      store_split4(ep, o + j, x);
    }
  }
}

// ---- coalescing epilogue (2-CTA kernel).  The accumulator layout is lane = row (32 rows per warp), so a direct
// store touches 32 different rows per instruction (32 partly filled sectors) and every epilogue load (bias, gamma,
// residual) queues behind the previous group's stores.  Here:
//  * the tile's bias / gamma slices (256 columns) are fetched into shared memory at the START of the tile, long before
//    the epilogue needs them (double buffered by tile parity, one named barrier per tile);
//  * a warp transposes 16 rows x 16 columns at a time through a 1 KB shared-memory tile (64 B per row, 16-byte chunks
//    XOR-swizzled by (row >> 1) & 3: conflict-free both ways), so that every store instruction writes 8 rows x 64
//    contiguous bytes (fp32) / 8 x 32 B (fp16 pairs): full sectors;
//  * LayerScale+residual in place (resid == out) is a vector reduction `red.global.add.v4.f32` -- one writer per
//    element, so the result is the same single fp32 addition, without the read round trip.
// Handles the row-major vector modes; returns false (nothing written) for the cases the direct path keeps (the
// transposed V third, ragged N / unaligned ldo).  Must be called by all 32 lanes.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// v[16]: raw accumulators of columns n..n+15 (PRE: final values of OUTPUT columns n..n+15, pair store only);
// sb / sg: shared-memory bias / gamma of those 16 columns.
template <bool PRE, int MODE>
__device__ __forceinline__ void epi_group16(const EpiParams& ep, const float* sb, const float* sg, float4* tile, int lane,
                                            int m_base, int n, int M, const float* v) {
  constexpr int mode = MODE;
  const int ch = lane & 3, nn = n + ch * 4;
  const float al = ep.alpha;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f), g = b;
  if (!PRE) b = *reinterpret_cast<const float4*>(sb + ch * 4);
  if (!PRE && mode == ANYLOC_EPI_LS_RESID) g = *reinterpret_cast<const float4*>(sg + ch * 4);
  const bool in_place = ep.resid == ep.out;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if ((lane >> 4) == h) {
      const int row = lane & 15, sw = (row >> 1) & 3;
#pragma unroll
      for (int c = 0; c < 4; ++c) tile[row * 4 + (c ^ sw)] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = i * 8 + (lane >> 2), m = m_base + h * 16 + r;
      const float4 a = tile[r * 4 + (ch ^ ((r >> 1) & 3))];
      if (m >= M) continue;
      const size_t o = (size_t)m * ep.ldo + nn;
      if (PRE) { store_split4(ep, o, a); continue; }
      float4 x = make_float4(a.x * al + b.x, a.y * al + b.y, a.z * al + b.z, a.w * al + b.w);
      if (mode == ANYLOC_EPI_BIAS) {
        *reinterpret_cast<float4*>(ep.out + o) = x;
      } else if (mode == ANYLOC_EPI_LS_RESID) {
        if (in_place) {
          red_add_v4(ep.out + o, make_float4(g.x * x.x, g.y * x.y, g.z * x.z, g.w * x.w));
        } else {
          const float4 rr = *reinterpret_cast<const float4*>(ep.resid + o);
          *reinterpret_cast<float4*>(ep.out + o) = make_float4(rr.x + g.x * x.x, rr.y + g.y * x.y, rr.z + g.z * x.z, rr.w + g.w * x.w);
        }
      } else if (mode == ANYLOC_EPI_QKV_SPLIT) {
        if (ep.qkv_f16) {
          uint2 hh, ll;
          split_f16x2(x.x * kActScale, x.y * kActScale, hh.x, ll.x);
          split_f16x2(x.z * kActScale, x.w * kActScale, hh.y, ll.y);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out) + o) = hh;
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out_lo) + o) = ll;
        } else {
          float4 hh, ll;
          split_tf32(x.x, hh.x, ll.x); split_tf32(x.y, hh.y, ll.y); split_tf32(x.z, hh.z, ll.z); split_tf32(x.w, hh.w, ll.w);
          *reinterpret_cast<float4*>(ep.out + o) = hh;
          *reinterpret_cast<float4*>(ep.out_lo + o) = ll;
        }
      } else {                               // BIAS_SPLIT / GELU_SPLIT
        if (mode == ANYLOC_EPI_GELU_SPLIT) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }
        store_split4(ep, o, x);
      }
    }
    __syncwarp();                            // the tile is rewritten by the next pass / call
  }
}

// 32 accumulator columns n..n+31 of one warp (lane = row m_base + lane); nl = n - (first column of the CTA tile).
// false = not handled here (caller falls back to the direct path).
template <int MODE>
__device__ __forceinline__ bool epi_chunk32_staged(const EpiParams& ep, const float* sbias, const float* sgamma,
                                                   float4* tile, int lane, int m_base, int n, int nl, int M, int N,
                                                   const float* v) {
  constexpr int mode = MODE;
  if (mode < 0) return true;                 // diagnostic: discard (ANYLOC_GEMM_DEBUG_SKIP_EPI)
  if (n + 32 > N || (ep.ldo & 3) || (mode == ANYLOC_EPI_QKV_SPLIT && n >= 2 * ep.qkv_D)) return false;
  if (mode == ANYLOC_EPI_SWIGLU_SPLIT) {
    // (x1_j, x2_j) interleaved -> 16 outputs silu(x1) * x2 at columns n/2.., activated in the lane = row layout
    // (bias reads are shared-memory broadcasts), then stored through the transposing tile
    const float al = ep.alpha;
    float y[16];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(sbias + nl + j);
      y[(j >> 1)] = silu_fast(v[j] * al + b.x) * (v[j + 1] * al + b.y);
      y[(j >> 1) + 1] = silu_fast(v[j + 2] * al + b.z) * (v[j + 3] * al + b.w);
    }
    epi_group16<true, MODE>(ep, nullptr, nullptr, tile, lane, m_base, n >> 1, M, y);
    return true;
  }
  epi_group16<false, MODE>(ep, sbias + nl, sgamma + nl, tile, lane, m_base, n, M, v);
  epi_group16<false, MODE>(ep, sbias + nl + 16, sgamma + nl + 16, tile, lane, m_base, n + 16, M, v + 16);
  return true;
}

// F16 = false: operands are fp32 words read as tf32 (32 elements per 128 B k-block, UMMA K=8, kind::tf32)
// F16 = true : operands are fp16            (64 elements per 128 B k-block, UMMA K=16, kind::f16, 2x rate)
template <int BN, bool F16, bool LO>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                int M, int N, int K, int has_a_lo, int has_b_lo, int band_n, EpiParams ep) {
  using C = Cfg<BN, LO>;
  if (ep.gate != nullptr && *reinterpret_cast<const volatile int*>(ep.gate) == 0) return;   // uniform over the grid
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* bar_area = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);          // [STAGES]
  uint64_t* empty_bar = full_bar + C::STAGES;                           // [STAGES]
  uint64_t* tfull_bar = empty_bar + C::STAGES;                          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  constexpr int BKE = F16 ? 64 : 32;         // elements per k-block (128 bytes)
  constexpr int CHUNK_KB = F16 ? CHUNK_KB_F16 : CHUNK_KB_TF32;
  const int num_k = (K + BKE - 1) / BKE;
  const int num_chunks = (num_k + CHUNK_KB - 1) / CHUNK_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
    if (has_a_lo) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
    if (has_b_lo) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // Register re-balancing inside the CTA's launch-time pool (640 x 96 = 61440):
  // 128 x 40 + 512 x 104 = 58368 <= 61440, so the blocking setmaxnreg.inc can always be satisfied.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------ TMA producer
      const uint32_t tx_bytes = A_BYTES * (1 + ((LO && has_a_lo) ? 1 : 0)) + C::B_BYTES * (1 + ((LO && has_b_lo) ? 1 : 0));
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb; tile_coords(tile, num_m, num_n, band_n, mb, nb);
        const int m0 = mb * BM, n0 = nb * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t fb = smem_u32(full_bar + stage);
          mbar_expect_tx(fb, tx_bytes);
          const uint32_t sbase = smem_u32(smem + stage * C::STAGE_BYTES);
          tma_load_2d(sbase, &tm_a_hi, fb, kb * BKE, m0);
          if (LO && has_a_lo) tma_load_2d(sbase + A_BYTES, &tm_a_lo, fb, kb * BKE, m0);
          tma_load_2d(sbase + C::B_OFF, &tm_b_hi, fb, kb * BKE, n0);
          if (LO && has_b_lo) tma_load_2d(sbase + C::B_OFF + C::B_BYTES, &tm_b_lo, fb, kb * BKE, n0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------ MMA issuer
      constexpr uint32_t fmt = F16 ? 0u : 2u;          // a/b format: 0 = F16, 2 = TF32; c format 1 = F32
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_k; ++kb) {
          const int in_chunk = kb % CHUNK_KB;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          if (in_chunk == 0) {                           // chunk accumulator must have been drained
            mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
            tc_fence_after();
          }
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_hi = make_desc(sbase), a_lo = make_desc(sbase + A_BYTES);
          const uint64_t b_hi = make_desc(sbase + C::B_OFF), b_lo = make_desc(sbase + C::B_OFF + C::B_BYTES);
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);      // +32 B per k-step inside the atom (both types)
            umma<F16>(d_tmem, a_hi + adv, b_hi + adv, idesc, (in_chunk | k) != 0);
            if (LO && has_a_lo) umma<F16>(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            if (LO && has_b_lo) umma<F16>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
          }
          umma_commit(smem_u32(empty_bar + stage));      // frees the smem stage when these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          if (in_chunk == CHUNK_KB - 1 || kb == num_k - 1) {
            umma_commit(smem_u32(tfull_bar + acc));      // chunk complete -> drain
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ---------------------------------------- accumulate (RN, registers) + epilogue: 16 warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int cq = (warp - 4) >> 2;          // column quarter: BN/4 columns
    constexpr int CPT = BN / 4;              // columns per thread
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb, nb; tile_coords(tile, num_m, num_n, band_n, mb, nb);
      const int m0 = mb * BM, n0 = nb * BN;
      float sum[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cq * CPT);
#pragma unroll
        for (int c = 0; c < CPT / 16; ++c) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(c * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[c * 16 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int m = m0 + q * 32 + lane;
      if (m < M) {
#pragma unroll
        for (int c = 0; c < CPT / 32; ++c) {
          const int n = n0 + cq * CPT + c * 32;
          if (n < N) epi_chunk32(ep, m, n, N, sum + c * 32);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ====================================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair (cluster of 2, same TPC) owns one 256 x 256 tile.  Each CTA loads ITS 128
// rows of A and ITS 128 of the 256 B rows (half the B bytes per SM -- the 1-CTA kernel needs ~62 B/clk/SM of L2->smem
// ingest, right at the per-SM limit), the leader CTA issues M256 x N256 UMMAs that read both CTAs' shared memory, and
// each CTA drains / post-processes its own 128 accumulator rows exactly like the 1-CTA kernel.  64 KB per stage ->
// 3-stage ring.  Barriers: `full` lives in the leader (armed once with the bytes of BOTH CTAs; both CTAs' TMA
// complete_tx on it), `empty` / `tfull` are signalled in both CTAs by multicast tcgen05.commit, `tempty` collects
// the drain arrivals of all 32 accumulate warps of the pair in the leader.
// ====================================================================================================================
namespace two {
constexpr int BH_BYTES = 128 * 128;                       // half of the B tile: 128 rows x 128 B
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;    // A_hi, A_lo, Bh_hi, Bh_lo = 64 KB
constexpr int STAGES = 3;
constexpr int HI_STAGE_BYTES = A_BYTES + BH_BYTES;         // hi-only pass: A_hi, Bh_hi = 32 KB
constexpr int HI_STAGES = 5;
constexpr int EPI_TILE_BYTES = 16 * 64;                    // per epilogue warp: 16 rows x 16 fp32 (coalescing transpose)
constexpr int EPI_VEC_BYTES = 2 * 2 * 256 * 4;              // bias + gamma slices of the tile's 256 columns, double buffered
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_VEC_BYTES + EPI_WARPS * EPI_TILE_BYTES;
static_assert(HI_STAGES * HI_STAGE_BYTES <= STAGES * STAGE_BYTES, "the hi-only ring must fit the same allocation");
constexpr int BN = 256;
}  // namespace two

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t local) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  // relaxed: the signal only says "this warp's tcgen05.ld of the buffer has completed" (tcgen05.wait::ld + fence
  // precede it); a release at cluster scope compiles to MEMBAR.ALL.GPU, which would also wait for the warp's
  // outstanding epilogue stores of the previous tile -- 18 % of the kernel's stall samples before this change
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (F16)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {       // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar) : "memory");
}

// LO = false: hi-only single pass (A_hi . B_hi^T, one MMA per product instead of three) -- the COARSE scores of the
// retrieval (topk.cu bounds their error rigorously and re-scores the candidates exactly).  Stages then hold
// {A_hi, Bh_hi} = 32 KB and the ring is HI_STAGES deep.  ep.gate (nullable): the kernel returns at once when *gate == 0.
template <bool F16, int MODE, bool LO = true>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc3_2cta_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                     const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                     int M, int N, int K, int band_n, int staged_epi, int chunk_kb, EpiParams ep) {
  using namespace two;
  if (ep.gate != nullptr && *reinterpret_cast<const volatile int*>(ep.gate) == 0) return;   // uniform over the grid
  constexpr int STAGES = LO ? two::STAGES : two::HI_STAGES;
  constexpr int STAGE_BYTES = LO ? two::STAGE_BYTES : two::HI_STAGE_BYTES;
  constexpr int A_LO_OFF = A_BYTES;                                      // LO only
  constexpr int B_HI_OFF = LO ? 2 * A_BYTES : A_BYTES;
  constexpr int B_LO_OFF = 2 * A_BYTES + BH_BYTES;                       // LO only
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* bar_area = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);          // [STAGES]  (used in the leader)
  uint64_t* empty_bar = full_bar + STAGES;                              // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                             // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                                 // [2]       (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();                 // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (M + 255) / 256, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  constexpr int BKE = F16 ? 64 : 32;
  const int CHUNK_KB = chunk_kb;             // k-blocks accumulated in TMEM between two round-to-nearest drains
  const int num_k = (K + BKE - 1) / BKE;
  const int num_chunks = (num_k + CHUNK_KB - 1) / CHUNK_KB;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(tfull_bar + s), 1); mbar_init(smem_u32(tempty_bar + s), 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // both CTAs' barriers are initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------ TMA producer (both CTAs; bytes land on the leader's barrier)
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int mb, nb; tile_coords(tile, num_m, num_n, band_n, mb, nb);
        const int m0 = mb * 256 + (int)rank * 128, n0 = nb * BN + (int)rank * 128;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t fb_local = smem_u32(full_bar + stage);
          if (rank == 0) mbar_expect_tx(fb_local, 2u * STAGE_BYTES);
          const uint32_t fb = mapa_rank0(fb_local);
          const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
          tma_load_2d_2sm(sbase, &tm_a_hi, fb, kb * BKE, m0);
          if (LO) tma_load_2d_2sm(sbase + A_LO_OFF, &tm_a_lo, fb, kb * BKE, m0);
          tma_load_2d_2sm(sbase + B_HI_OFF, &tm_b_hi, fb, kb * BKE, n0);
          if (LO) tma_load_2d_2sm(sbase + B_LO_OFF, &tm_b_lo, fb, kb * BKE, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      // ------------------------------------------------ MMA issuer (leader CTA only): M256 x N256 per instruction
      constexpr uint32_t fmt = F16 ? 0u : 2u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(256 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        for (int kb = 0; kb < num_k; ++kb) {
          const int in_chunk = kb % CHUNK_KB;
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          if (in_chunk == 0) {
            mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);
            tc_fence_after();
          }
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = make_desc(sbase), a_lo = make_desc(sbase + A_LO_OFF);
          const uint64_t b_hi = make_desc(sbase + B_HI_OFF), b_lo = make_desc(sbase + B_LO_OFF);
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma2<F16>(d_tmem, a_hi + adv, b_hi + adv, idesc, (in_chunk | k) != 0);
            if (LO) umma2<F16>(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            if (LO) umma2<F16>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
          }
          umma_commit_2sm(smem_u32(empty_bar + stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (in_chunk == CHUNK_KB - 1 || kb == num_k - 1) {
            umma_commit_2sm(smem_u32(tfull_bar + acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ---------------------------------------- accumulate (RN, registers) + epilogue: 16 warps per CTA, own 128 rows
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3;
    const int cq = (warp - 4) >> 2;
    constexpr int CPT = BN / 4;
    int acc = 0; uint32_t acc_phase = 0;
    int titer = 0;
    const uint32_t tempty_leader0 = mapa_rank0(smem_u32(tempty_bar)), tempty_leader1 = mapa_rank0(smem_u32(tempty_bar + 1));
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int mb, nb; tile_coords(tile, num_m, num_n, band_n, mb, nb);
      const int m0 = mb * 256 + (int)rank * 128, n0 = nb * BN;
      // bias / gamma slices of this tile -> shared memory now, while the first chunk is still being multiplied
      float* sbias = reinterpret_cast<float*>(bar_area + 256) + (titer & 1) * 512;
      float* sgamma = sbias + 256;
      {
        const int t = (int)threadIdx.x - 128;             // 0..511
        if (t < 128) {
          const bool isg = t >= 64;
          const float* src = isg ? ep.gamma : ep.bias;
          const int c = (t & 63) * 4;
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src) {
            if (n0 + c + 4 <= N) val = __ldg(reinterpret_cast<const float4*>(src + n0 + c));
            else {
              if (n0 + c < N) val.x = __ldg(src + n0 + c);
              if (n0 + c + 1 < N) val.y = __ldg(src + n0 + c + 1);
              if (n0 + c + 2 < N) val.z = __ldg(src + n0 + c + 2);
            }
          }
          *reinterpret_cast<float4*>((isg ? sgamma : sbias) + c) = val;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      ++titer;
      float sum[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cq * CPT);
#pragma unroll
        for (int c = 0; c < CPT / 16; ++c) {
          float v[16];
          tmem_ld16(trow + (uint32_t)(c * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[c * 16 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int m = m0 + q * 32 + lane;
      float4* etile = reinterpret_cast<float4*>(bar_area + 256 + EPI_VEC_BYTES + (warp - 4) * EPI_TILE_BYTES);
#pragma unroll
      for (int c = 0; c < CPT / 32; ++c) {
        const int nl = cq * CPT + c * 32, n = n0 + nl;
        if (n >= N) continue;                           // warp-uniform
        if (staged_epi && epi_chunk32_staged<MODE>(ep, sbias, sgamma, etile, lane, m0 + q * 32, n, nl, M, N, sum + c * 32))
          continue;
        if (m < M) epi_chunk32<MODE>(ep, m, n, N, sum + c * 32);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody exits (or frees TMEM) while the peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ host side
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

int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int ld, int box_rows, bool f16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("gemm_tc: cuTensorMapEncodeTiled unavailable"); return ANYLOC_ERR_CUDA; }
  const int esz = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld); return ANYLOC_ERR_CUDA; }
  return ANYLOC_OK;
}

}  // namespace tc

bool gemm_tc_supported(const void* a_hi, const void* a_lo, int lda, const void* b_hi, const void* b_lo, int ldb,
                       int M, int N, int K, const EpiParams& ep, bool f16) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int q = f16 ? 8 : 4;                 // elements per 16 bytes
  if (M < 1 || N < 1 || K < q) return false;
  if ((K % q) || (lda % q) || (ldb % q)) return false;
  if (!al16(a_hi) || !al16(b_hi) || (a_lo && !al16(a_lo)) || (b_lo && !al16(b_lo))) return false;
  if (!al16(ep.out) || (ep.out_lo && !al16(ep.out_lo)) || (ep.resid && !al16(ep.resid))) return false;
  if (ep.bias && !al16(ep.bias)) return false;
  if (ep.gamma && !al16(ep.gamma)) return false;
  return true;
}

template <bool F16>
static int launch_2cta(const void* a_hi, const void* a_lo, int lda, const void* b_hi, const void* b_lo, int ldb, int M,
                       int N, int K, const EpiParams& ep, int band_n, cudaStream_t st) {
  using namespace tc;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a_hi, M, K, lda, 128, F16))) return rc;
  if ((rc = make_map(&ma_lo, a_lo, M, K, lda, 128, F16))) return rc;
  if ((rc = make_map(&mb_hi, b_hi, N, K, ldb, 128, F16))) return rc;
  if ((rc = make_map(&mb_lo, b_lo, N, K, ldb, 128, F16))) return rc;
  const int tiles = cdiv(M, 256) * cdiv(N, two::BN);
  const int pairs = std::min(tiles, device_sm_count() / 2);
  static int staged_epi = -1;          // ANYLOC_GEMM_STAGED_EPI=0: direct (lane = row) stores, for A/B measurements
  if (staged_epi < 0) { const char* e = getenv("ANYLOC_GEMM_STAGED_EPI"); staged_epi = e ? atoi(e) : 1; }
  const int bn = std::min(band_n, cdiv(N, two::BN));
  // ANYLOC_GEMM_CHUNK: k-blocks per TMEM chunk (A/B knob; default 4 fp16 / 2 tf32 k-blocks = 48 / 24 MMAs per drain)
  static int chunk_env = -1;
  if (chunk_env < 0) { const char* e = getenv("ANYLOC_GEMM_CHUNK"); chunk_env = e ? atoi(e) : 0; }
  const int chunk = chunk_env > 0 ? chunk_env : (F16 ? CHUNK_KB_F16_2CTA : CHUNK_KB_TF32);
  // one instantiation per epilogue mode (compact per-tile code); -1 = the diagnostic "discard" variant
#define ANYLOC_LAUNCH_2CTA(MODE_)                                                                                   \
  case MODE_: {                                                                                                     \
    static unsigned long long attr_seen = 0;                                                                                   \
    if (first_use_on_this_device(&attr_seen)) {                                                                                                \
      ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc3_2cta_kernel<F16, MODE_>,                                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, two::SMEM_BYTES));        \
    }                                                                                                               \
    gemm_tc3_2cta_kernel<F16, MODE_><<<2 * pairs, THREADS, two::SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, M, N, \
                                                                                  K, bn, staged_epi, chunk, ep);   \
  } break;
  switch (ep.mode) {
    ANYLOC_LAUNCH_2CTA(-1)
    ANYLOC_LAUNCH_2CTA(ANYLOC_EPI_BIAS)
    ANYLOC_LAUNCH_2CTA(ANYLOC_EPI_BIAS_SPLIT)
    ANYLOC_LAUNCH_2CTA(ANYLOC_EPI_GELU_SPLIT)
    ANYLOC_LAUNCH_2CTA(ANYLOC_EPI_SWIGLU_SPLIT)
    ANYLOC_LAUNCH_2CTA(ANYLOC_EPI_LS_RESID)
    ANYLOC_LAUNCH_2CTA(ANYLOC_EPI_QKV_SPLIT)
    default: set_error("gemm_tc: unknown epilogue mode %d", ep.mode); return ANYLOC_ERR_ARG;
  }
#undef ANYLOC_LAUNCH_2CTA
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

// hi-only 2-CTA pass (fp16 operands, plain store epilogue): coarse retrieval scores
static int launch_2cta_hi(const void* a_hi, int lda, const void* b_hi, int ldb, int M, int N, int K, const EpiParams& ep,
                          int band_n, cudaStream_t st) {
  using namespace tc;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, a_hi, M, K, lda, 128, true))) return rc;
  if ((rc = make_map(&mb, b_hi, N, K, ldb, 128, true))) return rc;
  const int tiles = cdiv(M, 256) * cdiv(N, two::BN);
  const int pairs = std::min(tiles, device_sm_count() / 2);
  static unsigned long long attr_seen = 0;
  if (first_use_on_this_device(&attr_seen))
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc3_2cta_kernel<true, ANYLOC_EPI_BIAS, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, two::SMEM_BYTES));
  // 32 k-blocks (128 MMAs) per TMEM chunk: the coarse pass does not need the tight round-to-nearest accumulation
  gemm_tc3_2cta_kernel<true, ANYLOC_EPI_BIAS, false><<<2 * pairs, THREADS, two::SMEM_BYTES, st>>>(
      ma, ma, mb, mb, M, N, K, std::min(band_n, cdiv(N, two::BN)), 1, 32, ep);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

template <bool F16, bool LO>
static int launch_impl(const void* a_hi, const void* a_lo, int lda, const void* b_hi, const void* b_lo, int ldb, int M,
                       int N, int K, const EpiParams& ep, cudaStream_t st) {
  using namespace tc;
  constexpr int BN = 256;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a_hi, M, K, lda, BM, F16))) return rc;
  if ((rc = make_map(&ma_lo, a_lo ? a_lo : a_hi, M, K, lda, BM, F16))) return rc;
  if ((rc = make_map(&mb_hi, b_hi, N, K, ldb, BN, F16))) return rc;
  if ((rc = make_map(&mb_lo, b_lo ? b_lo : b_hi, N, K, ldb, BN, F16))) return rc;
  static unsigned long long attr_seen = 0;
  if (first_use_on_this_device(&attr_seen)) {
    ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc3_kernel<BN, F16, LO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<BN, LO>::SMEM_BYTES));
  }
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  const int grid = std::min(tiles, device_sm_count());
  static int band_n = -1;
  if (band_n < 0) { const char* e = getenv("ANYLOC_GEMM_BAND"); band_n = e ? atoi(e) : 8; if (band_n < 1) band_n = 1 << 20; }
  gemm_tc3_kernel<BN, F16, LO><<<grid, THREADS, Cfg<BN, LO>::SMEM_BYTES, st>>>(
      ma_hi, ma_lo, mb_hi, mb_lo, M, N, K, a_lo != nullptr, b_lo != nullptr, std::min(band_n, cdiv(N, BN)), ep);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

int gemm_tc_launch(const void* a_hi, const void* a_lo, int lda, const void* b_hi, const void* b_lo, int ldb, int M,
                   int N, int K, const EpiParams& ep_in, bool f16, cudaStream_t st) {
  const bool lo = a_lo != nullptr || b_lo != nullptr;
  static int band_n = -1, two_cta = -1;
  if (band_n < 0) { const char* e = getenv("ANYLOC_GEMM_BAND"); band_n = e ? atoi(e) : 8; if (band_n < 1) band_n = 1 << 20; }
  if (two_cta < 0) { const char* e = getenv("ANYLOC_GEMM_2CTA"); two_cta = e ? atoi(e) : 1; }
  // diagnostic only (tools/, never set by the product): bit m set -> GEMMs with epilogue mode m discard their result,
  // which exposes how much of a GEMM's time is its epilogue
  static int skip_epi = -1;
  if (skip_epi < 0) { const char* e = getenv("ANYLOC_GEMM_DEBUG_SKIP_EPI"); skip_epi = e ? atoi(e) : 0; }
  if (skip_epi && ((skip_epi >> ep_in.mode) & 1)) {
    EpiParams e2 = ep_in; e2.mode = -1;
    if (two_cta && a_lo && b_lo && M >= 512 && N >= 256)
      return f16 ? launch_2cta<true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, e2, band_n, st)
                 : launch_2cta<false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, e2, band_n, st);
  }
  const EpiParams& ep = ep_in;
  if (two_cta && a_lo && b_lo && M >= 512 && N >= 256) {
    return f16 ? launch_2cta<true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, ep, band_n, st)
               : launch_2cta<false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, ep, band_n, st);
  }
  if (two_cta && f16 && !lo && M >= 512 && N >= 256 && ep.mode == ANYLOC_EPI_BIAS)
    return launch_2cta_hi(a_hi, lda, b_hi, ldb, M, N, K, ep, band_n, st);
  if (f16) return lo ? launch_impl<true, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, ep, st)
                     : launch_impl<true, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, ep, st);
  return lo ? launch_impl<false, true>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, ep, st)
            : launch_impl<false, false>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, ep, st);
}

}  // namespace anyloc
