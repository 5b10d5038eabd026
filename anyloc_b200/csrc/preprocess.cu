// Image pre-processing on the device: the reference's `base_transform`
// (dvgl_benchmark/datasets_ws.py:20-23: ToTensor + Normalize(mean, std)) followed by the centre crop to a multiple of
// the patch size (scripts/dino_v2_vlad.py:174-176, demo/anyloc_vlad_generate.py:178-181), fused into one pass:
//   out[b,c,y,x] = ((float)img[b, top+y, left+x, c] / 255 - mean[c]) / std[c]
// Same operation order as torchvision (div, sub, div; IEEE round-to-nearest each) -> bit-identical results.
// HBM-bound: 3 B read + 12 B written per pixel.
#include <algorithm>
#include "common.cuh"

namespace anyloc {

__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int top, int left, int Hc, int Wc,
                     float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out) {
  // grid (ceil(Wc/2 / 256), Hc, B); a thread converts two neighbouring pixels and writes one float2 per plane
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= Wc) return;
  const uint8_t* src = img + (((size_t)b * H + top + y) * W + left + x) * 3;
  const bool two = x + 1 < Wc;
  float p[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) p[i] = (i < 3 || two) ? __fdiv_rn((float)__ldg(src + i), 255.0f) : 0.f;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  const size_t plane = (size_t)Hc * Wc;
  float* dst = out + (size_t)b * 3 * plane + (size_t)y * Wc + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float a = __fdiv_rn(__fsub_rn(p[c], mean[c]), sd[c]);
    float bb = __fdiv_rn(__fsub_rn(p[3 + c], mean[c]), sd[c]);
    if (two && ((Wc & 1) == 0)) *reinterpret_cast<float2*>(dst + c * plane) = make_float2(a, bb);
    else { dst[c * plane] = a; if (two) dst[c * plane + 1] = bb; }
  }
}


// ToTensor + Normalize + ANTIALIASED resize + centre crop in one pass (dvgl_benchmark/datasets_ws.py:222-239:
// `T.functional.resize(base_transform(img), [480, 640])`, bilinear; demo/anyloc_vlad_generate.py:165-177:
// `T.resize(img_pt, (h, w), InterpolationMode.BICUBIC)` of over-sized images).  On float tensors torchvision's resize is
// torch.nn.functional.interpolate(..., align_corners=False, antialias=True): per output index i,
//   scale = in / out; support = (taps/2) * max(scale, 1); centre = scale * (i + 0.5);
//   first = max(int(centre - support + 0.5), 0); n = min(int(centre + support + 0.5), in) - first;
//   w_j = filter((j + first - centre + 0.5) / max(scale, 1)), normalised to sum 1
// with the triangle filter (bilinear, 2 taps) or Keys' cubic with a = -0.5 (bicubic, 4 taps).  A thread owns one
// output pixel (all 3 channels): horizontal sums per source row, weighted by the vertical filter -- the order of the
// separable ATen CPU kernel (horizontal pass first).  Reads 3 B per tap, writes 12 B per pixel.
__device__ __forceinline__ float aa_filter(float x, int cubic) {
  x = fabsf(x);
  if (!cubic) return x < 1.0f ? 1.0f - x : 0.0f;
  const float a = -0.5f;
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}
struct AaSpan { int first, n; float scale_inv, centre; };
__device__ __forceinline__ AaSpan aa_span(int i, int in_size, int out_size, int cubic) {
  const float scale = (float)in_size / (float)out_size;
  const float support = (cubic ? 2.0f : 1.0f) * (scale >= 1.0f ? scale : 1.0f);
  AaSpan s;
  s.centre = scale * ((float)i + 0.5f);
  s.scale_inv = scale >= 1.0f ? 1.0f / scale : 1.0f;
  s.first = max((int)(s.centre - support + 0.5f), 0);
  s.n = min((int)(s.centre + support + 0.5f), in_size) - s.first;
  return s;
}
constexpr int AA_MAX_TAPS = 64;      // covers down-scaling by up to 16x (bicubic) / 32x (bilinear)

__global__ void __launch_bounds__(128)
preprocess_resize_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int Hr, int Wr, int cubic, int top, int left,
                            int Hc, int Wc, float m0, float m1, float m2, float s0, float s1, float s2,
                            float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= Wc) return;
  const AaSpan sx = aa_span(left + x, W, Wr, cubic), sy = aa_span(top + y, H, Hr, cubic);
  float wx[AA_MAX_TAPS];
  float totx = 0.f;
  for (int j = 0; j < sx.n; ++j) { wx[j] = aa_filter(((float)(j + sx.first) - sx.centre + 0.5f) * sx.scale_inv, cubic); totx += wx[j]; }
  for (int j = 0; j < sx.n; ++j) wx[j] /= totx;
  float toty = 0.f;
  for (int j = 0; j < sy.n; ++j) toty += aa_filter(((float)(j + sy.first) - sy.centre + 0.5f) * sy.scale_inv, cubic);
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  float acc[3] = {0.f, 0.f, 0.f};
  for (int jy = 0; jy < sy.n; ++jy) {
    const float wy = aa_filter(((float)(jy + sy.first) - sy.centre + 0.5f) * sy.scale_inv, cubic) / toty;
    const uint8_t* row = img + (((size_t)b * H + sy.first + jy) * W + sx.first) * 3;
    float h[3] = {0.f, 0.f, 0.f};
    for (int jx = 0; jx < sx.n; ++jx) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)__ldg(row + jx * 3 + c), 255.0f), mean[c]), sd[c]);
        h[c] = fmaf(wx[jx], v, h[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = fmaf(wy, h[c], acc[c]);
  }
  const size_t plane = (size_t)Hc * Wc;
  float* dst = out + (size_t)b * 3 * plane + (size_t)y * Wc + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[c * plane] = acc[c];
}

}  // namespace anyloc

using namespace anyloc;

extern "C" int anyloc_preprocess_u8(const uint8_t* img, int B, int H, int W, int top, int left, int Hc, int Wc,
                                    const float* mean3, const float* std3, float* out, void* stream) {
  ANYLOC_REQUIRE(img && out && mean3 && std3, "preprocess_u8: null pointer");
  ANYLOC_REQUIRE(B >= 0 && H > 0 && W > 0 && Hc > 0 && Wc > 0 && top >= 0 && left >= 0 && top + Hc <= H &&
                     left + Wc <= W,
                 "preprocess_u8: crop [%d+%d, %d+%d] outside the %dx%d image", top, Hc, left, Wc, H, W);
  ANYLOC_REQUIRE(Hc <= 65535 && B <= 65535, "preprocess_u8: Hc=%d / B=%d exceed the grid limits", Hc, B);
  ANYLOC_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "preprocess_u8: zero std");
  if (B == 0) return ANYLOC_OK;
  dim3 grid(cdiv(cdiv(Wc, 2), 256), Hc, B);
  preprocess_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, H, W, top, left, Hc, Wc, mean3[0], mean3[1], mean3[2],
                                                              std3[0], std3[1], std3[2], out);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}

// interpolation: 0 = bilinear, 1 = bicubic (both antialiased, torchvision's defaults for tensors).  The image is
// resized to Hr x Wr, then the window [top, top+Hc) x [left, left+Wc) of the RESIZED image is written.
extern "C" int anyloc_preprocess_resize_u8(const uint8_t* img, int B, int H, int W, int Hr, int Wr, int interpolation,
                                           int top, int left, int Hc, int Wc, const float* mean3, const float* std3,
                                           float* out, void* stream) {
  ANYLOC_REQUIRE(img && out && mean3 && std3, "preprocess_resize_u8: null pointer");
  ANYLOC_REQUIRE(interpolation == 0 || interpolation == 1, "preprocess_resize_u8: unknown interpolation %d", interpolation);
  ANYLOC_REQUIRE(B >= 0 && H > 0 && W > 0 && Hr > 0 && Wr > 0 && Hc > 0 && Wc > 0 && top >= 0 && left >= 0 &&
                     top + Hc <= Hr && left + Wc <= Wr,
                 "preprocess_resize_u8: crop [%d+%d, %d+%d] outside the resized %dx%d image", top, Hc, left, Wc, Hr, Wr);
  ANYLOC_REQUIRE(Hc <= 65535 && B <= 65535, "preprocess_resize_u8: Hc=%d / B=%d exceed the grid limits", Hc, B);
  ANYLOC_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "preprocess_resize_u8: zero std");
  const float taps = interpolation ? 4.0f : 2.0f;
  const float sxm = std::max((float)W / Wr, 1.0f), sym = std::max((float)H / Hr, 1.0f);
  ANYLOC_REQUIRE(taps * sxm + 2.0f <= AA_MAX_TAPS && taps * sym + 2.0f <= 1.0e9f,
                 "preprocess_resize_u8: horizontal down-scaling factor %.1f exceeds the %d-tap window", sxm, AA_MAX_TAPS);
  if (B == 0) return ANYLOC_OK;
  dim3 grid(cdiv(Wc, 128), Hc, B);
  preprocess_resize_u8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(img, H, W, Hr, Wr, interpolation, top, left, Hc, Wc,
                                                                      mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                                                      std3[2], out);
  ANYLOC_CHECK_LAUNCH();
  return ANYLOC_OK;
}
