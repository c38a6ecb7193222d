// Hardware probe (B200, sm_100a): steady-state cost of one tcgen05.mma (M=128, K=16, bf16, SS mode,
// cta_group::1) as a function of N, with the issue loop on the uniform datapath (elect.sync, unrolled),
// for the un-swizzled "channel-planar halo" A operand of csrc/conv3d.cu.  The question it answers: is
// a narrow-N MMA bound by the tensor pipe (128*N/256 cycles) or by the shared-memory operand fetch
// (A: 4 KB + B: N*32 B per MMA)?  Output is recorded in profiles/.  Not product code.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/probe_mma_rate tools/probe_mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../vdm4cdm_b200/csrc/ptx.cuh"

using namespace vdm;

// mode 0: every MMA reads the same A and B addresses; mode 1: A start address walks over the 27 tap
// offsets x 4 slices of a (6,18,10) halo, B walks over 8 weight blocks (the conv kernel's pattern);
// mode 2: like 1 but A is a dense SWIZZLE_NONE tile (SBO = 128 B, no halo pitch).
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int trips, int mode, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x * 16; i < 200 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) {
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint32_t a16 = ptx::smem_u32(smem) >> 4, b16 = a16 + (96 * 1024 >> 4);
    const uint32_t P16 = 6 * 18 * 10;   // halo plane, 16-byte units
    const uint32_t sbo = mode == 2 ? 8u : 10u;
    const uint64_t hi_a = ((uint64_t)(P16 & 0x3FFF) << 16) | ((uint64_t)sbo << 32) | ((uint64_t)1 << 46);
    const uint64_t hi_b = ((uint64_t)((uint32_t)N & 0x3FFF) << 16) | ((uint64_t)8 << 32) | ((uint64_t)1 << 46);
    const bool leader = ptx::elect_one();
    long long t0 = 0;
    if (leader) {
      t0 = clock64();
      for (int t = 0; t < trips; ++t) {
        const uint32_t tap = (uint32_t)t % 27u;
        const uint32_t a_tap = mode == 0 ? a16 : a16 + ((tap / 9u) * 180u + ((tap / 3u) % 3u) * 10u + tap % 3u);
        const uint32_t b_tap = mode == 0 ? b16 : b16 + (tap & 7u) * (uint32_t)(4 * N);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t sa = mode == 0 ? 0u : (uint32_t)s * 180u + (uint32_t)(2 * j) * P16;
            const uint32_t sb = mode == 0 ? 0u : (uint32_t)(2 * j) * (uint32_t)N;
            const uint64_t ad = hi_a | (uint64_t)((a_tap + sa) & 0x3FFFu);
            const uint64_t bd = hi_b | (uint64_t)((b_tap + sb) & 0x3FFFu);
            ptx::umma_bf16(tmem_base + (uint32_t)((s * N) % 512 + N <= 512 ? (s * N) % 512 : 0), ad, bd, idesc, 1u);
          }
        }
      }
      ptx::umma_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    if (leader) cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base_s, 512);
  }
}

template <int N>
static void run(int mode, int grid, long long* d_cyc) {
  const int trips = 2000;
  cudaError_t e = cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) printf("cudaFuncSetAttribute: %s\n", cudaGetErrorString(e));
  cudaMemset(d_cyc, 0xff, sizeof(long long) * 148);
  rate_kernel<N><<<grid, 128, 210 * 1024>>>(trips, mode, d_cyc);
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("launch: %s\n", cudaGetErrorString(e));
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    exit(2);
  }
  long long h[148];
  cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mn = h[0], mx = h[0];
  for (int i = 1; i < grid; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
  const double per = (double)mx / (trips * 8.0), per_min = (double)mn / (trips * 8.0);
  const double floor_cyc = 128.0 * N / 256.0, smem_cyc = (4096.0 + N * 32.0) / 128.0;
  printf("N=%3d mode=%d grid=%3d: %.1f cyc/MMA (fastest CTA %.1f); tensor floor %.0f, smem model (A+B)/128B %.0f -> %.0f%% of tensor peak\n",
         N, mode, grid, per, per_min, floor_cyc, smem_cyc, 100.0 * floor_cyc / per);
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, sizeof(long long) * 148);
  for (int grid : {1, 148})
    for (int mode = 0; mode < 3; ++mode) {
      run<32>(mode, grid, d_cyc);
      run<64>(mode, grid, d_cyc);
      run<96>(mode, grid, d_cyc);
      run<128>(mode, grid, d_cyc);
      run<192>(mode, grid, d_cyc);
      run<256>(mode, grid, d_cyc);
    }
  printf("probe done\n");
  return 0;
}
