// Read-bandwidth microbenchmark for the access patterns of the VLAD kernels (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/membw tools/membw.cu && gpurun_out/membw
// Patterns over a [R, D] fp32 matrix (R = 16928 x 4 rows, D = 1536: 416 MB, larger than L2):
//   seq      : every CTA streams a contiguous chunk, warp = 512 contiguous bytes per load, U loads in flight
//   slice512 : CTA = (128-column slice, 529-row image): 512-byte pieces at a 6 KB stride (accumulate kernels)
//   box128   : CTA = 128 consecutive rows, sweeps D in 128-byte k-blocks: one warp instruction = 4 rows x 128 B
//              (what a [128 x 32 float] TMA box asks of DRAM)
//   box512   : same tile, 512 bytes of a row per warp instruction
//   copy     : read + write (the MEASURED_PEAKS.json definition), for reference
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int U>
__global__ void __launch_bounds__(256) seq_kernel(const float4* __restrict__ x, size_t n4, float* out) {
  float acc = 0.f;
  const size_t per_cta = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t b0 = (size_t)blockIdx.x * per_cta, b1 = min(n4, b0 + per_cta);
  for (size_t i = b0 + threadIdx.x; i < b1; i += 256 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { size_t j = i + (size_t)u * 256; v[u] = j < b1 ? __ldg(x + j) : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256) slice512_kernel(const float* __restrict__ x, int N, int D, float* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* xb = x + (size_t)blockIdx.y * N * D + blockIdx.x * 128 + lane * 4;
  float acc = 0.f;
  for (int n0 = w * U; n0 < N; n0 += 8 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { int n = min(n0 + u, N - 1); v[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)n * D)); }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

// slice512 with an indirection: rows visited in the order perm[] (per image), optionally rotated per slice
template <int U>
__global__ void __launch_bounds__(256) slice512_perm_kernel(const float* __restrict__ x, const int* __restrict__ perm, int N, int D,
                                                            int rotate, float* out) {
  extern __shared__ int dyn[];             // only to limit the CTAs per SM
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* xb = x + (size_t)blockIdx.y * N * D + blockIdx.x * 128 + lane * 4;
  const int* pb = perm + (size_t)blockIdx.y * N;
  const int rot = rotate ? (int)(((long long)blockIdx.x * N) / gridDim.x) : 0;
  float acc = 0.f;
  for (int n0 = w * U; n0 < N; n0 += 8 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { int n = n0 + u; n = n < N ? n : N - 1; n += rot; n = n >= N ? n - N : n;
      v[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)pb[n] * D)); }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[0] = acc + dyn[0];
}

// CTA = (J*128-column slice, image): each warp reads J*512 contiguous bytes of a row (J loads per lane), U rows in flight
template <int J, int U>
__global__ void __launch_bounds__(256) slicew_kernel(const float* __restrict__ x, int N, int D, float* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* xb = x + (size_t)blockIdx.y * N * D + blockIdx.x * 128 * J + lane * 4;
  float acc = 0.f;
  for (int n0 = w * U; n0 < N; n0 += 8 * U) {
    float4 v[U][J];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int n = min(n0 + u, N - 1);
#pragma unroll
      for (int j = 0; j < J; ++j) v[u][j] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)n * D + j * 128));
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < J; ++j) acc += v[u][j].x + v[u][j].y + v[u][j].z + v[u][j].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

// CTA = 128 rows; warp w owns rows [16w, 16w+16); BYTES contiguous bytes of a row per warp instruction part
template <int BYTES, int U>
__global__ void __launch_bounds__(256) box_kernel(const float* __restrict__ x, int R, int D, float* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int LPR = BYTES / 16;            // lanes per row piece
  constexpr int RPI = 32 / LPR;              // rows per warp instruction
  float acc = 0.f;
  for (int tile = blockIdx.x; tile * 128 < R; tile += gridDim.x) {
    const int r0 = tile * 128 + w * 16;
    for (int c0 = 0; c0 < D * 4; c0 += BYTES) {                  // sweep D
      // 16 rows per warp: 16 / RPI instructions, all in flight (U unused when 16/RPI is small)
      float4 v[16 / RPI];
#pragma unroll
      for (int i = 0; i < 16 / RPI; ++i) {
        const int r = min(r0 + i * RPI + lane / LPR, R - 1);
        v[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(x + (size_t)r * D) + c0 + (lane % LPR) * 16));
      }
#pragma unroll
      for (int i = 0; i < 16 / RPI; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) y[i] = __ldg(x + i);
}

template <typename F> float time_it(F f, int iters = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  const int N = 529, D = 1536, B = 128, R = B * N;
  const size_t n = (size_t)R * D, bytes = n * 4;
  float *x, *y, *out;
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&y, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(x, 0, bytes));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  auto rep = [&](const char* name, float ms, double b) { printf("%-28s %8.1f us  %7.0f GB/s\n", name, ms * 1e3, b / ms / 1e6); fflush(stdout); };
  rep("copy (read+write bytes)", time_it([&] { copy_kernel<<<sms * 8, 256>>>((const float4*)x, (float4*)y, n / 4); }), 2.0 * bytes);
  rep("seq U=4, 8 CTAs/SM", time_it([&] { seq_kernel<4><<<sms * 8, 256>>>((const float4*)x, n / 4, out); }), bytes);
  rep("seq U=8, 8 CTAs/SM", time_it([&] { seq_kernel<8><<<sms * 8, 256>>>((const float4*)x, n / 4, out); }), bytes);
  rep("seq U=8, 4 CTAs/SM", time_it([&] { seq_kernel<8><<<sms * 4, 256>>>((const float4*)x, n / 4, out); }), bytes);
  rep("seq U=16, 4 CTAs/SM", time_it([&] { seq_kernel<16><<<sms * 4, 256>>>((const float4*)x, n / 4, out); }), bytes);
  rep("slice512 U=4", time_it([&] { slice512_kernel<4><<<dim3(D / 128, B), 256>>>(x, N, D, out); }), bytes);
  rep("slice512 U=8", time_it([&] { slice512_kernel<8><<<dim3(D / 128, B), 256>>>(x, N, D, out); }), bytes);
  rep("slice512 U=16", time_it([&] { slice512_kernel<16><<<dim3(D / 128, B), 256>>>(x, N, D, out); }), bytes);
  {
    // identity, label-sorted (32 clusters, random labels), and fully random row orders
    int* perm; CK(cudaMalloc(&perm, (size_t)B * N * 4));
    int* h = (int*)malloc((size_t)B * N * 4);
    auto run = [&](const char* name, int rotate, size_t dsm) {
      CK(cudaMemcpy(perm, h, (size_t)B * N * 4, cudaMemcpyHostToDevice));
      CK(cudaFuncSetAttribute(slice512_perm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      rep(name, time_it([&] { slice512_perm_kernel<4><<<dim3(D / 128, B), 256, dsm>>>(x, perm, N, D, rotate, out); }), bytes);
    };
    for (int b = 0; b < B; ++b) for (int n = 0; n < N; ++n) h[b * N + n] = n;
    run("perm identity U=4", 0, 0);
    run("perm identity, 3 CTAs/SM", 0, 70 * 1024);
    run("perm identity, 2 CTAs/SM", 0, 100 * 1024);
    run("perm identity, rotated/slice", 1, 0);
    srand(1);
    for (int b = 0; b < B; ++b) {          // stable sort by random label (32 clusters)
      int lab[529], pos = 0;
      for (int n = 0; n < N; ++n) lab[n] = rand() % 32;
      for (int k = 0; k < 32; ++k) for (int n = 0; n < N; ++n) if (lab[n] == k) h[b * N + pos++] = n;
    }
    run("perm label-sorted", 0, 0);
    run("perm label-sorted, 3 CTAs/SM", 0, 70 * 1024);
    run("perm label-sorted, rotated", 1, 0);
    for (int b = 0; b < B; ++b) for (int n = N - 1; n > 0; --n) { int j = rand() % (n + 1); int t = h[b * N + n]; h[b * N + n] = h[b * N + j]; h[b * N + j] = t; }
    run("perm random", 0, 0);
    run("perm random, rotated", 1, 0);
  }
  rep("slice 2 KB (J=4) U=2", time_it([&] { slicew_kernel<4, 2><<<dim3(D / 512, B), 256>>>(x, N, D, out); }), bytes);
  rep("slice 2 KB (J=4) U=4", time_it([&] { slicew_kernel<4, 4><<<dim3(D / 512, B), 256>>>(x, N, D, out); }), bytes);
  rep("row 6 KB (J=12) U=1", time_it([&] { slicew_kernel<12, 1><<<dim3(1, B), 256>>>(x, N, D, out); }), bytes);
  rep("row 6 KB (J=12) U=2", time_it([&] { slicew_kernel<12, 2><<<dim3(1, B), 256>>>(x, N, D, out); }), bytes);
  rep("box128 (4 rows x 128 B)", time_it([&] { box_kernel<128, 1><<<sms * 2, 256>>>(x, R, D, out); }), bytes);
  rep("box256 (2 rows x 256 B)", time_it([&] { box_kernel<256, 1><<<sms * 2, 256>>>(x, R, D, out); }), bytes);
  rep("box512 (1 row x 512 B)", time_it([&] { box_kernel<512, 1><<<sms * 2, 256>>>(x, R, D, out); }), bytes);
  rep("box512, 4 CTAs/SM", time_it([&] { box_kernel<512, 1><<<sms * 4, 256>>>(x, R, D, out); }), bytes);
  return 0;
}
