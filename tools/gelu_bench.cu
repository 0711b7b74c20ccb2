// Microbenchmark: throughput of GELU formulations on one SM-filling grid (elements / clk / SM), to find what bounds the GELU epilogues
// of the fc1 / fused MLP kernels.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gelu_bench tools/gelu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vit-ocm-wmsegmentation_b200/csrc/gemm_sm100.cuh"
using namespace vitocm;

__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__device__ __forceinline__ void gelu32(float (&v)[32]) {
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_sigmoid5_x2(v[j], v[j + 1]);
  } else if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) gelu_sigmoid_x2(v[j], v[j + 1]);
  } else if (MODE == 2) {   // scalar five-coefficient form
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = v[j], u = fminf(x * x, 30.f);
      float w = fmaf(u, -3.2290010e-06f, 8.8238336e-05f);
      w = fmaf(w, u, 3.6027357e-04f); w = fmaf(w, u, -0.10522669f); w = fmaf(w, u, -2.3020453f);
      v[j] = x * ptx::rcp_approx(1.0f + ptx::ex2_approx(x * w));
    }
  } else if (MODE == 3) {   // the two MUFU per element only
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ptx::rcp_approx(ptx::ex2_approx(v[j]));
  } else if (MODE == 4) {   // the FMA-pipe part of mode 0 only
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const uint64_t x2 = ptx::pack_f32x2(v[j], v[j + 1]);
      const uint64_t sq = ptx::mul_f32x2(x2, x2);
      float u0, u1; ptx::unpack_f32x2(sq, u0, u1);
      const uint64_t u = ptx::pack_f32x2(fminf(u0, 30.f), fminf(u1, 30.f));
      uint64_t w = ptx::fma_f32x2(u, ptx::dup_f32x2(-3.2290010e-06f), ptx::dup_f32x2(8.8238336e-05f));
      w = ptx::fma_f32x2(w, u, ptx::dup_f32x2(3.6027357e-04f));
      w = ptx::fma_f32x2(w, u, ptx::dup_f32x2(-0.10522669f));
      w = ptx::fma_f32x2(w, u, ptx::dup_f32x2(-2.3020453f));
      const uint64_t arg = ptx::mul_f32x2(x2, w);
      const uint64_t den = ptx::add_f32x2(arg, ptx::dup_f32x2(1.0f));
      const uint64_t r = ptx::mul_f32x2(x2, den);
      ptx::unpack_f32x2(r, v[j], v[j + 1]);
    }
  } else if (MODE == 5) {   // one MUFU per element (tanh form; speed reference only -- 2^-11 relative error)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = v[j], u = fminf(x * x, 30.f);
      float w = fmaf(u, -3.2290010e-06f, 8.8238336e-05f);
      w = fmaf(w, u, 3.6027357e-04f); w = fmaf(w, u, -0.10522669f); w = fmaf(w, u, -2.3020453f);
      const float hx = 0.5f * x;
      v[j] = fmaf(hx, tanh_approx(x * w), hx);
    }
  } else if (MODE == 6) {   // one MUFU only
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ptx::ex2_approx(v[j]);
  } else if (MODE == 7) {   // rcp only
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ptx::rcp_approx(v[j]);
  } else if (MODE == 8) {   // pure FMA-pipe Phi(x) polynomial stand-in: 11 packed FMAs per pair, no MUFU
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const uint64_t x2 = ptx::pack_f32x2(v[j], v[j + 1]);
      const uint64_t sq = ptx::mul_f32x2(x2, x2);
      float u0, u1; ptx::unpack_f32x2(sq, u0, u1);
      const uint64_t u = ptx::pack_f32x2(fminf(u0, 25.f), fminf(u1, 25.f));
      uint64_t w = ptx::dup_f32x2(1.0e-9f);
#pragma unroll
      for (int k = 0; k < 11; ++k) w = ptx::fma_f32x2(w, u, ptx::dup_f32x2(0.001f * (k + 1)));
      const uint64_t r = ptx::mul_f32x2(x2, ptx::fma_f32x2(x2, w, ptx::dup_f32x2(0.5f)));
      ptx::unpack_f32x2(r, v[j], v[j + 1]);
    }
  }
}

template <int MODE>
__global__ void k(float* out, const float* in, int iters, long long* clk) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = in[(threadIdx.x + j * 17) & 1023];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    gelu32<MODE>(v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] * 0.5f + 0.3f;   // keep the values in range (1 extra FFMA per element, all modes)
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE> void run(const char* name, int threads) {
  float *out, *in; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 1024 * 4); cudaMalloc(&clk, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = -3.f + 6.f * i / 1024.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, in, 10, clk);
  k<MODE><<<148, threads>>>(out, in, iters, clk);
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %2d warps/SM: %6.2f elements/clk/SM  (%lld clk; %s)\n", name, threads / 32, 32.0 * threads * iters / c, c, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(in); cudaFree(clk);
}
int main() {
  for (int threads : {256, 512}) {
    run<0>("sigmoid5 packed (fp16 engines)", threads);
    run<1>("sigmoid3 packed (bf16 engines)", threads);
    run<2>("sigmoid5 scalar", threads);
    run<3>("ex2 + rcp only", threads);
    run<4>("FMA-pipe part of sigmoid5 packed only", threads);
    run<5>("sigmoid5 scalar, tanh.approx (1 MUFU)", threads);
    run<6>("ex2 only", threads);
    run<7>("rcp only", threads);
    run<8>("11 packed FMAs per pair, no MUFU", threads);
  }
  return 0;
}
