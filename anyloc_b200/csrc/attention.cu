// Multi-head self-attention of the DINOv2 block (upstream dinov2/layers/attention.py, reached from
// /root/reference/utilities.py:269): softmax(q k^T / 8) v per head (head_dim 64), fp32.
// v1: SIMT flash-style kernel, one CTA per (64-query tile, head, image); online softmax;
// output written as the tf32 (hi,lo) pair that feeds the `proj` GEMM.
#include "common.cuh"

namespace anyloc {

constexpr int AT = 64;          // query tile = key tile = head_dim
constexpr int ATP = AT + 4;     // padded row

__global__ void __launch_bounds__(256)
attention_simt_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_lo, int T, int D,
                      float* __restrict__ o_hi, float* __restrict__ o_lo, int out_f16) {
  extern __shared__ float sm[];
  float* Qt = sm;                 // [d][row]
  float* Kt = Qt + AT * ATP;      // [d][key]
  float* Vs = Kt + AT * ATP;      // [key][d]
  float* Pt = Vs + AT * ATP;      // [key][row]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * AT, h = blockIdx.y, b = blockIdx.z;
  const size_t ld = (size_t)3 * D;
  const float* base = qkv + (size_t)b * T * ld + (size_t)h * AT;
  const float* base_lo = qkv_lo ? qkv_lo + (size_t)b * T * ld + (size_t)h * AT : nullptr;
  auto ld4 = [&](size_t off) {
    float4 v = __ldg(reinterpret_cast<const float4*>(base + off));
    if (base_lo) {
      float4 w = __ldg(reinterpret_cast<const float4*>(base_lo + off));
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    return v;
  };
  const int lrow = tid >> 4, lc = (tid & 15) * 4;

  // Q tile (transposed)
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int r = lrow + p * 16, t = q0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) v = ld4((size_t)t * ld + lc);
    Qt[(lc + 0) * ATP + r] = v.x; Qt[(lc + 1) * ATP + r] = v.y;
    Qt[(lc + 2) * ATP + r] = v.z; Qt[(lc + 3) * ATP + r] = v.w;
  }
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f; }

  for (int k0 = 0; k0 < T; k0 += AT) {
    __syncthreads();   // previous tile's Kt/Vs/Pt fully consumed (also covers the Q stores)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int r = lrow + p * 16, t = k0 + r;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (t < T) {
        kv = ld4((size_t)t * ld + D + lc);
        vv = ld4((size_t)t * ld + 2 * D + lc);
      }
      Kt[(lc + 0) * ATP + r] = kv.x; Kt[(lc + 1) * ATP + r] = kv.y;
      Kt[(lc + 2) * ATP + r] = kv.z; Kt[(lc + 3) * ATP + r] = kv.w;
      *reinterpret_cast<float4*>(&Vs[r * ATP + lc]) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < AT; ++d) {
      float4 a = *reinterpret_cast<const float4*>(&Qt[d * ATP + ty * 4]);
      float4 c = *reinterpret_cast<const float4*>(&Kt[d * ATP + tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], cv[j], s[i][j]);
    }
    // scale, mask, online softmax (row statistics across the 16 tx lanes)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx * 4 + j < T) ? s[i][j] * 0.125f : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(m[i], mx);          // finite: every tile has >= 1 valid key
      const float corr = expf(m[i] - mnew);         // exp(-inf) = 0 on the first tile
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - mnew); rs += s[i][j]; }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l[i] = l[i] * corr + rs;
      m[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) { o[i][j] *= corr; Pt[(tx * 4 + j) * ATP + ty * 4 + i] = s[i][j]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < AT; ++k) {
      float4 p = *reinterpret_cast<const float4*>(&Pt[k * ATP + ty * 4]);
      float4 v = *reinterpret_cast<const float4*>(&Vs[k * ATP + tx * 4]);
      float pv[4] = {p.x, p.y, p.z, p.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pv[i], vv[j], o[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int t = q0 + ty * 4 + i;
    if (t >= T) continue;
    float inv = 1.0f / l[i];
    size_t off = ((size_t)b * T + t) * D + (size_t)h * AT + tx * 4;
    if (out_f16) {
      __half hh[4], ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_f16(o[i][j] * inv * kActScale, hh[j], ll[j]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(o_hi) + off) = *reinterpret_cast<uint2*>(hh);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(o_lo) + off) = *reinterpret_cast<uint2*>(ll);
    } else {
      float4 hh, ll;
      split_tf32(o[i][0] * inv, hh.x, ll.x); split_tf32(o[i][1] * inv, hh.y, ll.y);
      split_tf32(o[i][2] * inv, hh.z, ll.z); split_tf32(o[i][3] * inv, hh.w, ll.w);
      *reinterpret_cast<float4*>(o_hi + off) = hh;
      *reinterpret_cast<float4*>(o_lo + off) = ll;
    }
  }
}

int attention_launch(const float* qkv, const float* qkv_lo, int B, int T, int D, int heads, void* o_hi,
                     void* o_lo, bool out_f16, cudaStream_t st) {
  ANYLOC_REQUIRE(D == heads * AT, "attention: head_dim must be 64 (D=%d heads=%d)", D, heads);
  size_t smem = (size_t)4 * AT * ATP * sizeof(float);
  ANYLOC_CHECK_CUDA(cudaFuncSetAttribute(attention_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  attention_simt_kernel<<<dim3(cdiv(T, AT), heads, B), 256, smem, st>>>(qkv, qkv_lo, T, D, (float*)o_hi, (float*)o_lo, out_f16 ? 1 : 0);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

}  // namespace anyloc
