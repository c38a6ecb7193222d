// Special-function-unit throughput on sm_100a: warp-instructions of tanh / ex2 / rcp per clock per SM, measured with 8
// independent chains per thread, 1..8 warps per scheduler.   nvcc -arch=sm_100a -O3 -o tools/bin/probe_mufu tools/probe_mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) { unsigned u = __float_as_uint(x), v; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v); }
  if (OP == 4) { unsigned u = __float_as_uint(x), v; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v); }
  if (OP == 5) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 6) { unsigned u = __float_as_uint(x), v; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(v) : "r"(u)); y = __uint_as_float(v); }
  if (OP == 7) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[8];
  for (int j = 0; j < 8; ++j) v[j] = 0.1f * (threadIdx.x + j + 1);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = op<OP>(v[j]);
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 4096;
  for (int threads : {128, 256, 512, 1024}) {
    k<OP><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double warp_instr = (double)iters * 8 * (threads / 32);
    printf("%-22s %4d threads/SM: %.2f cycles per warp-instruction per SM (%.2f lanes/clk/SM)\n", name, threads, c / warp_instr,
           32.0 * warp_instr / c);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tanh.approx.f32"); run<1>("ex2.approx.ftz.f32"); run<2>("rcp.approx.ftz.f32"); run<5>("rsqrt.approx.ftz.f32");
  run<3>("tanh.approx.f16x2"); run<4>("tanh.approx.bf16x2"); run<6>("ex2.approx.ftz.bf16x2"); run<7>("fma.rn.f32");
  return 0;
}
