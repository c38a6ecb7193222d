// TMA tiled-load throughput vs. box row length (no swizzle, 4-D planar tensor [planes][D][H][W*8 bf16]).
// Each CTA (one per SM) repeatedly loads a box (row_elems, n_h, n_d, n_planes) at marching coordinates into a
// 2-stage shared-memory ring and reports cycles per load; rows = n_h*n_d*n_planes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/probe_tma tools/probe_tma.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void probe(const __grid_constant__ CUtensorMap tm, int box_bytes, int iters, int W, int H, int D, int xoff, int stages, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int stage_bytes = (box_bytes + 1023) & ~1023;
    long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      if (i >= stages) mbar_wait(&bar[(i - stages) % stages], ((i - stages) / stages) & 1);
      if (i < iters) {
        const int s = i % stages;
        mbar_expect(&bar[s], box_bytes);
        const int t = blockIdx.x * 131 + i;
        tma4(smem + s * stage_bytes, &tm, &bar[s], ((t % (W / 8)) * 8) * 8 + xoff * 8, ((t / 16) % (H / 16)) * 16, (t / 128) % (D - 8), 0);
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int W = 128, H = 128, D = 128, P = 8;
  const size_t elems = (size_t)P * D * H * W * 8;
  void* dptr;
  cudaMalloc(&dptr, elems * 2);
  cudaMemset(dptr, 0, elems * 2);
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  PFN_cuTensorMapEncodeTiled_v12000 enc;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int wv, nh, nd, np, xoff; const char* name; };
  Cfg cfgs[] = {
      {10, 18, 6, 4, -1, "conv halo 10x18x6 x4 planes (160 B rows, start -16 B)"},
      {10, 18, 6, 4, 0, "same, 128-B aligned start"},
      {8, 18, 6, 4, 0, "8 voxels (128 B rows, aligned)"},
      {18, 10, 6, 4, -1, "18x10x6 x4 planes (288 B rows)"},
      {18, 18, 6, 2, -1, "18x18x6 x2 planes (288 B rows)"},
      {16, 10, 6, 4, 0, "16 voxels (256 B rows, aligned)"},
      {34, 10, 3, 4, -1, "34x10x3 x4 planes (544 B rows)"},
      {10, 18, 1, 4, -1, "one slice 10x18 x4 planes"},
      {10, 18, 4, 2, -1, "10x18x4 x2 planes (MT=2, KC=16)"},
      {8, 16, 1, 4, 0, "wgrad g tile 8x16 x4 planes"},
  };
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)P};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)c.wv * 8, (cuuint32_t)c.nh, (cuuint32_t)c.nd, (cuuint32_t)c.np};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, c.name); continue; }
    const int box_bytes = c.wv * 16 * c.nh * c.nd * c.np;
    const int rows = c.nh * c.nd * c.np;
    for (int stages = 1; stages <= 2; ++stages) {
      const int stage_bytes = (box_bytes + 1023) & ~1023;
      if (stages * stage_bytes > 200 * 1024) continue;
      const int iters = 200;
      probe<<<148, 32, stages * stage_bytes>>>(tm, box_bytes, iters, W, H, D, c.xoff, stages, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      avg /= 148.0 * iters;
      printf("%-56s stages=%d: %8.0f cyc/load  %6.1f cyc/row  %6.2f B/clk/SM  (%d rows, %d B)\n", c.name, stages, avg, avg / rows,
             box_bytes / avg, rows, box_bytes);
    }
  }
  return 0;
}
