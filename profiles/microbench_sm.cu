// SM-level throughput probes that the attention kernel's design rests on (B200, one CTA on one SM):
//   * tcgen05.ld 32x32b.x32 bytes/clk with 4 and 8 warps (is the TMEM read port per SM or per SM sub-partition?)
//   * MUFU.EX2 (f32) / MUFU.EX2.BF16 warp instructions per clk per sub-partition
//   * FFMA2 (fma.rn.f32x2) and FFMA issue rate
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/bin/microbench_sm profiles/microbench_sm.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dove_b200/csrc/ptx.cuh"
using namespace dove;

__global__ void tmem_ld_probe(long long* out, int iters, int chunks) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[128];
    for (int c = 0; c < 4; ++c)
      if (c < chunks) tmem_ld32(base + c * 32, v + c * 32);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < chunks) acc += v[c * 32] ^ v[c * 32 + 31];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  if (warp == 0) tmem_dealloc<512>(tptr);
}

template <int MODE>
__global__ void alu_probe(long long* out, float* sink, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  uint64_t y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %1};" : "=l"(y[i]) : "f"(x[i]));
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = 0xbc00bc00u + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(y[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (MODE == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(y[i]));
    s += x[i] + a + b + __uint_as_float(h[i]);
  }
  if (s == 1234.5f) sink[0] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 16);
  cudaMalloc(&sink, 16);
  long long h[2];
  const int iters = 2000;
  for (int warps : {4, 8}) {
    for (int chunks : {1, 4}) {
      tmem_ld_probe<<<1, warps * 32, 0>>>(d, iters, chunks);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * warps * chunks * 32 * 32 * 4;
      printf("tcgen05.ld 32x32b.x32 x%d per wait, %d warps: %lld clk, %.1f B/clk/SM, %.1f clk per warp-instruction\n",
             chunks, warps, h[0], bytes / h[0], double(h[0]) / (double(iters) * chunks));
    }
  }
  const char* names[5] = {"MUFU.EX2 f32", "ex2.bf16x2 (2 MUFU.EX2.BF16 + PRMT)", "FFMA2 (f32x2)", "FFMA", "ex2.f16x2"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 5; ++mode) {
      if (mode == 0) alu_probe<0><<<1, warps * 32>>>(d, sink, iters);
      if (mode == 1) alu_probe<1><<<1, warps * 32>>>(d, sink, iters);
      if (mode == 2) alu_probe<2><<<1, warps * 32>>>(d, sink, iters);
      if (mode == 3) alu_probe<3><<<1, warps * 32>>>(d, sink, iters);
      if (mode == 4) alu_probe<4><<<1, warps * 32>>>(d, sink, iters);
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      const double insts = double(iters) * 8 * warps;
      printf("%-38s %2d warps: %.2f clk per warp-instruction per sub-partition (%.1f lane-results/clk/SM)\n", names[mode],
             warps, double(h[0]) / (insts / 4), insts * 32 * ((mode == 1 || mode == 2 || mode == 4) ? 2 : 1) / h[0]);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
