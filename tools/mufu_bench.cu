// Microbenchmark: MUFU.EX2 throughput for f32 vs packed bf16x2 / f16x2 operands (one SM-filling grid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + blockIdx.x, b = a ^ 0x9e3779b9u, c = a + 12345u, d = b + 999u;
  float fa = __uint_as_float((a & 0x007fffffu) | 0xbf000000u), fb = fa * 0.9f, fc = fa * 0.8f, fd = fa * 0.7f;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fa)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fb));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fc)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(fd));
      fa -= 1.5f; fb -= 1.5f; fc -= 1.5f; fd -= 1.5f;
    } else if (MODE == 1) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(c)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(d));
      a ^= 0x80008000u; b ^= 0x80008000u; c ^= 0x80008000u; d ^= 0x80008000u;
    } else {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
      a ^= 0x80008000u; b ^= 0x80008000u; c ^= 0x80008000u; d ^= 0x80008000u;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ __float_as_uint(fa + fb + fc + fd);
}
template <int MODE> void run(const char* name, int values_per_instr) {
  uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(out, 100);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = 148.0 * 8 * 256 * 4.0 * iters;
  printf("%s: %.3f ms, %.1f G thread-instr/s, %.1f G values/s (err %s)\n", name, ms, instr / ms / 1e6, instr * values_per_instr / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() { run<0>("ex2.f32   ", 1); run<1>("ex2.bf16x2", 2); run<2>("ex2.f16x2 ", 2); return 0; }
