// Sibling aggregators on the same patch features: GeM (scripts/dino_v2_gem.py:170-189) and
// average / max pooling (scripts/dino_v2_gp.py:130-135) over the patch axis of [B,N,D] features -> [B,D].
//   average : mean_n x                      max : max_n x
//   gem     : m = mean_n x^p (|x|^p with use_abs);  out = sign(m) |m|^(1/p)   (the reference takes the complex
//             root and restores the sign; with use_abs it is the plain real root)
// One read of the features (HBM-bound, 4 B per element).  CTA = (128-column slice, image): 8 row groups x 32 lanes
// x float4 columns, fixed-order shared-memory reduction across the row groups -> deterministic.
#include "common.cuh"

namespace anyloc {

enum { POOL_AVG = 0, POOL_MAX = 1, POOL_GEM = 2 };

__device__ __forceinline__ float gem_pow(float x, float p, int ip, bool use_abs) {
  if (use_abs) x = fabsf(x);
  if (ip > 0) {                       // integer exponent: repeated products like torch.pow(x, 3)
    float r = x;
    for (int i = 1; i < ip; ++i) r *= x;
    return r;
  }
  return powf(x, p);                  // NaN for negative x and fractional p, as in the reference
}

// torch.max propagates NaN (fmaxf drops it)
__device__ __forceinline__ float nanmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }

template <int MODE>
__global__ void __launch_bounds__(256)
pool_kernel(const float* __restrict__ x, const int32_t* __restrict__ n_valid, int N, int D, float p, int ip,
            int use_abs, float* __restrict__ out) {
  __shared__ float4 part[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int b = blockIdx.y, col = blockIdx.x * 128 + lane * 4;
  const int n = n_valid ? min(N, n_valid[b]) : N;
  const bool colok = col < D;
  float4 acc = MODE == POOL_MAX ? make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (colok) {
    const float* xb = x + (size_t)b * N * D + col;
    for (int r = grp; r < n; r += 8) {
      float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)r * D));
      if (MODE == POOL_MAX) {
        acc.x = nanmax(acc.x, v.x); acc.y = nanmax(acc.y, v.y); acc.z = nanmax(acc.z, v.z); acc.w = nanmax(acc.w, v.w);
      } else if (MODE == POOL_GEM) {
        acc.x += gem_pow(v.x, p, ip, use_abs); acc.y += gem_pow(v.y, p, ip, use_abs);
        acc.z += gem_pow(v.z, p, ip, use_abs); acc.w += gem_pow(v.w, p, ip, use_abs);
      } else {
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  part[grp][lane] = acc;
  __syncthreads();
  if (grp == 0 && colok) {
    float r[4] = {acc.x, acc.y, acc.z, acc.w};
    for (int g = 1; g < 8; ++g) {
      float4 o = part[g][lane];
      float q[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (MODE == POOL_MAX) r[i] = nanmax(r[i], q[i]);
        else r[i] += q[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE != POOL_MAX) r[i] = r[i] / (float)n;
      if (MODE == POOL_GEM) {
        const float m = r[i];
        const float root = powf(fabsf(m), 1.0f / p);
        r[i] = use_abs ? root : (m > 0.f ? root : (m < 0.f ? -root : (m == 0.f ? 0.f : m)));
      }
    }
    *reinterpret_cast<float4*>(out + (size_t)b * D + col) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

}  // namespace anyloc

using namespace anyloc;

extern "C" int anyloc_pool(const float* feats, const int32_t* n_valid, int B, int N, int D, int mode, float gem_p,
                           int gem_use_abs, float* out, void* stream) {
  ANYLOC_REQUIRE(feats && out, "pool: null pointer");
  ANYLOC_REQUIRE(B >= 0 && N > 0 && D > 0 && D % 4 == 0, "pool: bad dims B=%d N=%d D=%d (D multiple of 4)", B, N, D);
  ANYLOC_REQUIRE(B <= 65535, "pool: B=%d exceeds the grid limit", B);
  ANYLOC_REQUIRE(mode >= POOL_AVG && mode <= POOL_GEM, "pool: unknown mode %d", mode);
  ANYLOC_REQUIRE(mode != POOL_GEM || gem_p != 0.f, "pool: gem_p must be non-zero");
  if (B == 0) return ANYLOC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(D, 128), B);
  const int ip = (gem_p == floorf(gem_p) && gem_p >= 1.f && gem_p <= 16.f) ? (int)gem_p : 0;
  if (mode == POOL_AVG) pool_kernel<POOL_AVG><<<grid, 256, 0, st>>>(feats, n_valid, N, D, gem_p, ip, gem_use_abs, out);
  else if (mode == POOL_MAX) pool_kernel<POOL_MAX><<<grid, 256, 0, st>>>(feats, n_valid, N, D, gem_p, ip, gem_use_abs, out);
  else pool_kernel<POOL_GEM><<<grid, 256, 0, st>>>(feats, n_valid, N, D, gem_p, ip, gem_use_abs, out);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}
