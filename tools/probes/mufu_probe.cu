// Throughput of the MUFU exp2 forms on sm_100a: f32, f16x2, bf16x2 (two results per instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu_probe.bin tools/probes/mufu_probe.cu && ./mufu_probe.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void probe(uint32_t* out, long long* cycles, int iters) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0xbc00bc00u + threadIdx.x + i;      // small negative halves / bf16s
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) f[i] = ex2_f32(f[i]) - 1.0f;
      else if (MODE == 1) a[i] = ex2_f16x2(a[i]) ^ 0x80008000u;
      else a[i] = ex2_bf16x2(a[i]) ^ 0x80008000u;
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= a[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char* names[3] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2"};
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int mode = 0; mode < 3; ++mode) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<148, warps * 32>>>(out, cyc, iters);
        else if (mode == 1) probe<1><<<148, warps * 32>>>(out, cyc, iters);
        else probe<2><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double instr_per_sm = double(iters) * 8 * warps;               // warp-instructions
      printf("%-24s warps/SM %2d: %.2f cycles per warp-instruction per SM (%.1f results/clk/SM)\n", names[mode], warps,
             double(h) / instr_per_sm, instr_per_sm * 32 * (mode ? 2 : 1) / double(h));
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
