// Image pre-processing on the device: the reference's `base_transform`
// (dvgl_benchmark/datasets_ws.py:20-23: ToTensor + Normalize(mean, std)) followed by the centre crop to a multiple of
// the patch size (scripts/dino_v2_vlad.py:174-176, demo/anyloc_vlad_generate.py:178-181), fused into one pass:
//   out[b,c,y,x] = ((float)img[b, top+y, left+x, c] / 255 - mean[c]) / std[c]
// Same operation order as torchvision (div, sub, div; IEEE round-to-nearest each) -> bit-identical results.
// HBM-bound: 3 B read + 12 B written per pixel.
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
