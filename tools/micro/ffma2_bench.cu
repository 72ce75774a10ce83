// Microbenchmark: FFMA / FFMA2 issue throughput per SM sub-partition on B200, and LDS.64/LDS.128 issue cost.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0.001f * i;
  __syncthreads();
  float2 acc[5][4];
  for (int p = 0; p < 5; ++p) for (int c = 0; c < 4; ++c) acc[p][c] = make_float2(0.f, 0.f);
  float2 w[5]; float4 xv = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int p = 0; p < 5; ++p) w[p] = make_float2(0.5f + p, 0.25f + p);
  const float* wp = sm + (threadIdx.x / 8) * 10;
  const float* xp = sm + 2048 + (threadIdx.x % 8) * 4;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1 || MODE == 3) {  // with shared loads (same pattern as the MLP tile)
      xv = *reinterpret_cast<const float4*>(xp + (it & 15) * 32);
#pragma unroll
      for (int p = 0; p < 5; ++p) w[p] = *reinterpret_cast<const float2*>(wp + (it & 15) * 50 + 2 * p);
    }
    if (MODE == 0 || MODE == 1) {
      const float2 x0 = make_float2(xv.x, xv.x), x1 = make_float2(xv.y, xv.y), x2 = make_float2(xv.z, xv.z), x3 = make_float2(xv.w, xv.w);
#pragma unroll
      for (int p = 0; p < 5; ++p) {
        acc[p][0] = __ffma2_rn(w[p], x0, acc[p][0]);
        acc[p][1] = __ffma2_rn(w[p], x1, acc[p][1]);
        acc[p][2] = __ffma2_rn(w[p], x2, acc[p][2]);
        acc[p][3] = __ffma2_rn(w[p], x3, acc[p][3]);
      }
    } else {  // scalar FFMA, same math
#pragma unroll
      for (int p = 0; p < 5; ++p) {
        acc[p][0].x = fmaf(w[p].x, xv.x, acc[p][0].x); acc[p][0].y = fmaf(w[p].y, xv.x, acc[p][0].y);
        acc[p][1].x = fmaf(w[p].x, xv.y, acc[p][1].x); acc[p][1].y = fmaf(w[p].y, xv.y, acc[p][1].y);
        acc[p][2].x = fmaf(w[p].x, xv.z, acc[p][2].x); acc[p][2].y = fmaf(w[p].y, xv.z, acc[p][2].y);
        acc[p][3].x = fmaf(w[p].x, xv.w, acc[p][3].x); acc[p][3].y = fmaf(w[p].y, xv.w, acc[p][3].y);
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int p = 0; p < 5; ++p) for (int c = 0; c < 4; ++c) s += acc[p][c].x + acc[p][c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char* names[4] = {"FFMA2 regs only", "FFMA2 + LDS(1x128,5x64)", "FFMA regs only", "FFMA + LDS"};
  for (int mode = 0; mode < 4; ++mode)
    for (int nt = 128; nt <= 1024; nt *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, nt, 32768>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, nt, 32768>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, nt, 32768>>>(out, cyc, iters);
        if (mode == 3) k<3><<<148, nt, 32768>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_iter = (double)h / iters;  // cycles per iteration of 20 FFMA2 (=40 FMA lanes-instr) per warp
      const int wps = nt / 32 / 4;
      printf("%-26s warps/SMSP=%d: %.1f cycles/iter/warp -> %.2f cycles per FFMA2-equivalent per SMSP, %.1f FMA/clk/SM\n", names[mode], wps,
             per_iter, per_iter / (20.0 * wps), 4.0 * wps * 20 * 64 / per_iter);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
