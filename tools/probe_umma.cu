// Hardware probe (B200, sm_100a): which shared-memory operand layouts tcgen05.mma accepts for
// "row-shifted views of a halo tile", and how fast narrow-N MMAs issue.  Not product code: its
// output (profiles/r01_probe_umma.txt) records the facts the conv kernels in
// vdm4cdm_b200/csrc/ are designed around.
//
//   T1  K-major, no swizzle, channel-planar halo ([c/8][h'][w'][8ch]); tap shift = start-address
//       offset of (dh*Wh+dw)*16 B.  Both (LBO,SBO) role assignments are tried.
//   T2  K-major SWIZZLE_128B rows of 64 bf16 (NDHWC halo), row-shifted start, base_offset 0 / phase.
//   T3  K-major SWIZZLE_64B rows of 32 bf16, row-shifted start.
//   T4  MN-major, no swizzle (the wgrad operands: channels along M/N, voxels along K).
//   T5  issue rate: cycles per M=128,K=16 MMA for N in {16..256}, planar vs SW128 operands.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/probe_umma tools/probe_umma.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../vdm4cdm_b200/csrc/ptx.cuh"

using namespace vdm;

constexpr int kMaxMma = 64;

struct ProbeParams {
  int a_bytes, b_bytes;
  int n;
  int n_mma;
  int repeat;
  uint32_t idesc;
  uint64_t a_desc_hi, b_desc_hi;  // every field except the start address
  int a_off[kMaxMma], b_off[kMaxMma];
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, ProbeParams p,
             float* __restrict__ d_out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((p.a_bytes + 1023) & ~1023);
  for (int i = threadIdx.x * 16; i < p.a_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(a_s + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (int i = threadIdx.x * 16; i < p.b_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(b_s + i) = *reinterpret_cast<const uint4*>(b_img + i);
  ptx::fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  int cols = 32;
  while (cols < p.n) cols <<= 1;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_s, (uint32_t)cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t a_base = ptx::smem_u32(a_s), b_base = ptx::smem_u32(b_s);
    const long long t0 = clock64();
    for (int r = 0; r < p.repeat; ++r) {
      for (int i = 0; i < p.n_mma; ++i) {
        const uint64_t ad = p.a_desc_hi | (uint64_t)(((a_base + (uint32_t)p.a_off[i]) & 0x3FFFFu) >> 4);
        const uint64_t bd = p.b_desc_hi | (uint64_t)(((b_base + (uint32_t)p.b_off[i]) & 0x3FFFFu) >> 4);
        ptx::umma_bf16(tmem_base, ad, bd, p.idesc, (r > 0 || i > 0) ? 1u : 0u);
      }
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (cycles) *cycles = t1 - t0;
  }
  __syncthreads();
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < p.n; c0 += 16) {
    uint32_t raw[16];
    ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, raw);
    ptx::tmem_ld_wait();
    const int m = warp * 32 + (threadIdx.x & 31);
    for (int j = 0; j < 16; ++j) d_out[m * p.n + c0 + j] = __uint_as_float(raw[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)cols);
  }
}

// T6: tensor-pipe rate with the issue loop out of the way: the same (A, B) descriptors, 16 MMAs per trip.
// T7: the conv kernel's issue pattern: per-MMA descriptor arithmetic (tap offsets from a table in smem).
__global__ void __launch_bounds__(128, 1)
rate_kernel(int n, int trips, int mode, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int tap_off[32];
  for (int i = threadIdx.x * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 27) tap_off[threadIdx.x] = ((threadIdx.x / 9) * 180 + ((threadIdx.x / 3) % 3) * 10 + threadIdx.x % 3) * 16;
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
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = ptx::smem_u32(smem), b_base = a_base + 96 * 1024;
    const uint32_t P = 6 * 18 * 10 * 16;
    const uint64_t hi_a = ((uint64_t)((P >> 4) & 0x3FFF) << 16) | ((uint64_t)((160 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    const uint64_t hi_b = ((uint64_t)(((uint32_t)n * 16 >> 4) & 0x3FFF) << 16) | ((uint64_t)((128 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
    const long long t0 = clock64();
    if (mode == 0) {
      const uint64_t ad = hi_a | (uint64_t)((a_base & 0x3FFFFu) >> 4), bd = hi_b | (uint64_t)((b_base & 0x3FFFFu) >> 4);
      for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int i = 0; i < 16; ++i) ptx::umma_bf16(tmem_base, ad, bd, idesc, 1u);
      }
    } else {
      // 27 taps x 4 sub-tiles x 2 k16 = 216 MMAs per trip, descriptors rebuilt per MMA like conv3d.cu
      for (int t = 0; t < trips; ++t) {
        for (int tap = 0; tap < 27; ++tap) {
          const uint32_t a_tap = a_base + (uint32_t)tap_off[tap];
          const uint32_t b_tap = b_base + (uint32_t)(tap & 7) * (uint32_t)(4 * n * 16);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t ad = hi_a | (uint64_t)(((a_tap + (uint32_t)s * 2880u + (uint32_t)(2 * j) * P) & 0x3FFFFu) >> 4);
              const uint64_t bd = hi_b | (uint64_t)(((b_tap + (uint32_t)(2 * j) * (uint32_t)n * 16u) & 0x3FFFFu) >> 4);
              ptx::umma_bf16(tmem_base + (uint32_t)(s * n), ad, bd, idesc, 1u);
            }
          }
        }
      }
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    *cycles = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host ---------------------------------------------------------------------------------------
static uint64_t make_desc_hi(uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}
static uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
static int ival(int a, int b, int c, int seed) {  // small integers in [-3, 3]
  uint32_t h = ((uint32_t)a * 73856093u) ^ ((uint32_t)b * 19349663u) ^ ((uint32_t)c * 83492791u) ^ ((uint32_t)seed * 2654435761u);
  h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
  return (int)(h % 7u) - 3;
}
static void put(std::vector<uint8_t>& img, size_t off, int v) {
  __nv_bfloat16 h = __float2bfloat16((float)v);
  memcpy(&img[off], &h, 2);
}

static uint8_t *d_a, *d_b;
static float* d_d;
static long long* d_cyc;

static bool run(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, ProbeParams& p, std::vector<float>& out,
                long long* cyc) {
  p.a_bytes = (int)a.size();
  p.b_bytes = (int)b.size();
  cudaMemcpy(d_a, a.data(), a.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice);
  cudaMemset(d_d, 0xff, 128 * 256 * sizeof(float));
  const size_t smem = ((a.size() + 1023) & ~(size_t)1023) + ((b.size() + 1023) & ~(size_t)1023) + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  probe_kernel<<<1, 128, smem>>>(d_a, d_b, p, d_d, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  CUDA error: %s\n", cudaGetErrorString(e));
    return false;
  }
  out.resize(128 * p.n);
  cudaMemcpy(out.data(), d_d, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
  if (cyc) cudaMemcpy(cyc, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  return true;
}

static int compare(const std::vector<float>& got, const std::vector<float>& want, int n) {
  int bad = 0;
  for (int i = 0; i < 128 * n; ++i)
    if (got[i] != want[i]) ++bad;
  return bad;
}

constexpr int HH = 18, WH = 10;  // halo rows / row pitch in voxels (16h x 8w tile + 1 voxel border)

// T1: planar no-swizzle K-major
static void test_planar(int n, int ktot) {
  const int planes = ktot / 8;
  const int P = HH * WH * 16;
  std::vector<uint8_t> a((size_t)planes * P, 0), b((size_t)planes * n * 16, 0);
  for (int k = 0; k < ktot; ++k)
    for (int h = 0; h < HH; ++h)
      for (int w = 0; w < WH; ++w) put(a, (size_t)(k / 8) * P + (h * WH + w) * 16 + (k % 8) * 2, ival(k, h, w, 1));
  for (int k = 0; k < ktot; ++k)
    for (int j = 0; j < n; ++j) put(b, (size_t)(k / 8) * n * 16 + j * 16 + (k % 8) * 2, ival(k, j, 0, 2));
  const int shifts[5][2] = {{0, 0}, {1, 1}, {2, 2}, {1, 0}, {0, 1}};
  for (int hyp = 0; hyp < 1; ++hyp) {  // hyp 1 (roles swapped) faults on hardware: addresses leave shared memory
    for (auto& s : shifts) {
      ProbeParams p;
      memset(&p, 0, sizeof(p));
      p.n = n; p.repeat = 1; p.n_mma = ktot / 16;
      p.idesc = make_idesc(128, n, 0, 0);
      // hyp 0: LBO = K-direction core-matrix stride, SBO = 8-row-group stride.  hyp 1: swapped.
      const uint32_t a_k = P, a_m = WH * 16, b_k = n * 16, b_m = 128;
      p.a_desc_hi = hyp == 0 ? make_desc_hi(a_k, a_m, 0, 0) : make_desc_hi(a_m, a_k, 0, 0);
      p.b_desc_hi = hyp == 0 ? make_desc_hi(b_k, b_m, 0, 0) : make_desc_hi(b_m, b_k, 0, 0);
      for (int j = 0; j < p.n_mma; ++j) {
        p.a_off[j] = (s[0] * WH + s[1]) * 16 + j * 2 * P;
        p.b_off[j] = j * 2 * n * 16;
      }
      std::vector<float> want(128 * n), got;
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < n; ++j) {
          float acc = 0;
          for (int k = 0; k < ktot; ++k) acc += (float)ival(k, m / 8 + s[0], m % 8 + s[1], 1) * (float)ival(k, j, 0, 2);
          want[m * n + j] = acc;
        }
      if (!run(a, b, p, got, nullptr)) exit(2);
      printf("T1 planar-noswizzle N=%d K=%d hyp=%s shift=(%d,%d): mismatches=%d\n", n, ktot,
             hyp == 0 ? "LBO=K,SBO=MN" : "LBO=MN,SBO=K", s[0], s[1], compare(got, want, n));
    }
  }
}

// T2/T3: swizzled NDHWC halo rows; row_bytes = 128 (SW128) or 64 (SW64)
static void test_swizzled(int n, int row_bytes, int wh) {
  const int ktot = row_bytes / 2;
  const int chunks = row_bytes / 16;            // 16-byte chunks per row
  const int layout = row_bytes == 128 ? 2 : 4;
  const int rows = HH * wh;
  std::vector<uint8_t> a((size_t)rows * row_bytes, 0), b((size_t)n * row_bytes, 0);
  auto swz = [&](int row, int k) {
    const size_t row_off = (size_t)row * row_bytes;
    const int phase = (int)((row_off >> 7) & (chunks - 1));  // address bits [7,..) select the XOR phase
    return row_off + (size_t)(((k / 8) ^ phase) * 16) + (k % 8) * 2;
  };
  for (int h = 0; h < HH; ++h)
    for (int w = 0; w < wh; ++w)
      for (int k = 0; k < ktot; ++k) put(a, swz(h * wh + w, k), ival(k, h, w, 3));
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < ktot; ++k) put(b, swz(j, k), ival(k, j, 0, 4));
  const int shifts[5][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}, {2, 2}};
  for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
    for (auto& s : shifts) {
      ProbeParams p;
      memset(&p, 0, sizeof(p));
      p.n = n; p.repeat = 1; p.n_mma = ktot / 16;
      p.idesc = make_idesc(128, n, 0, 0);
      const int start = (s[0] * wh + s[1]) * row_bytes;
      const uint32_t bo = bo_mode ? (uint32_t)((start >> 7) & 7) : 0u;
      p.a_desc_hi = make_desc_hi(16, (uint32_t)(wh * row_bytes), layout, bo);
      p.b_desc_hi = make_desc_hi(16, (uint32_t)(8 * row_bytes), layout, 0);
      for (int j = 0; j < p.n_mma; ++j) {
        p.a_off[j] = start + j * 32;
        p.b_off[j] = j * 32;
      }
      std::vector<float> want(128 * n), got;
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < n; ++j) {
          float acc = 0;
          for (int k = 0; k < ktot; ++k) acc += (float)ival(k, m / 8 + s[0], m % 8 + s[1], 3) * (float)ival(k, j, 0, 4);
          want[m * n + j] = acc;
        }
      if (!run(a, b, p, got, nullptr)) exit(2);
      printf("T%d SW%d rowpitch=%d voxels base_offset=%s shift=(%d,%d): mismatches=%d\n", row_bytes == 128 ? 2 : 3,
             row_bytes, wh, bo_mode ? "phase" : "0", s[0], s[1], compare(got, want, n));
    }
  }
}

// T4: MN-major no-swizzle operands (wgrad): A[m=ci][k=voxel], B[n=co][k=voxel]; 8 channels x 16 B per voxel.
static void test_mn_major(int n) {
  const int m_planes = 16, n_planes = n / 8;  // M = 128 channels
  const int vox = 2 * WH;                      // two halo rows of WH voxels; K = 16 voxels = 8 from each row
  const int PA = vox * 16, PB = 16 * 16;      // plane strides (B is dense: 16 voxels)
  std::vector<uint8_t> a((size_t)m_planes * PA, 0), b((size_t)n_planes * PB, 0);
  for (int c = 0; c < 128; ++c)
    for (int v = 0; v < vox; ++v) put(a, (size_t)(c / 8) * PA + v * 16 + (c % 8) * 2, ival(c, v, 0, 5));
  for (int c = 0; c < n; ++c)
    for (int v = 0; v < 16; ++v) put(b, (size_t)(c / 8) * PB + v * 16 + (c % 8) * 2, ival(c, v, 0, 6));
  for (int hyp = 0; hyp < 1; ++hyp) {
    for (int dw = 0; dw < 3; ++dw) {
      ProbeParams p;
      memset(&p, 0, sizeof(p));
      p.n = n; p.repeat = 1; p.n_mma = 1;
      p.idesc = make_idesc(128, n, 1, 1);
      // hyp 0: SBO = stride between 8-channel groups (MN), LBO = stride between 8-voxel groups (K)
      const uint32_t a_mn = PA, a_k = WH * 16, b_mn = PB, b_k = 128;
      p.a_desc_hi = hyp == 0 ? make_desc_hi(a_k, a_mn, 0, 0) : make_desc_hi(a_mn, a_k, 0, 0);
      p.b_desc_hi = hyp == 0 ? make_desc_hi(b_k, b_mn, 0, 0) : make_desc_hi(b_mn, b_k, 0, 0);
      p.a_off[0] = dw * 16;
      p.b_off[0] = 0;
      std::vector<float> want(128 * n), got;
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < n; ++j) {
          float acc = 0;
          for (int k = 0; k < 16; ++k) {
            const int v = (k / 8) * WH + (k % 8) + dw;
            acc += (float)ival(m, v, 0, 5) * (float)ival(j, k, 0, 6);
          }
          want[m * n + j] = acc;
        }
      if (!run(a, b, p, got, nullptr)) exit(2);
      printf("T4 MN-major-noswizzle N=%d hyp=%s dw=%d: mismatches=%d\n", n, hyp == 0 ? "LBO=K,SBO=MN" : "LBO=MN,SBO=K", dw,
             compare(got, want, n));
    }
  }
}

// T5: issue rate
static void test_rate(int n, int mode) {  // mode 0: planar no-swizzle (K=32 halo); 1: SW128 rows; 2: SW64 rows
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.repeat = 64; p.n_mma = 54;  // 27 taps x K=32
  p.idesc = make_idesc(128, n, 0, 0);
  std::vector<uint8_t> a, b;
  if (mode == 0) {
    const int P = HH * WH * 16;
    a.assign((size_t)4 * P, 0);
    const int tb = std::min(27, (144 * 1024) / (4 * n * 16));  // distinct weight taps that fit in smem
    b.assign((size_t)tb * 4 * n * 16, 0);
    p.a_desc_hi = make_desc_hi(P, WH * 16, 0, 0);
    p.b_desc_hi = make_desc_hi(n * 16, 128, 0, 0);
    for (int t = 0; t < 27; ++t)
      for (int j = 0; j < 2; ++j) {
        const int dh = (t / 3) % 3, dw = t % 3;
        p.a_off[t * 2 + j] = (dh * WH + dw) * 16 + j * 2 * P;
        p.b_off[t * 2 + j] = ((t % tb) * 4 + j * 2) * n * 16;
      }
  } else {
    const int rb = mode == 1 ? 128 : 64;
    a.assign((size_t)HH * 16 * rb, 0);
    const int tb = std::min(27, (144 * 1024) / (n * 64));
    b.assign((size_t)tb * n * 64, 0);  // taps x [n][32 ch]
    p.a_desc_hi = make_desc_hi(16, 16 * rb, mode == 1 ? 2 : 4, 0);
    p.b_desc_hi = make_desc_hi(16, 8 * 64, 4, 0);
    for (int t = 0; t < 27; ++t)
      for (int j = 0; j < 2; ++j) {
        const int dh = (t / 3) % 3;
        p.a_off[t * 2 + j] = dh * 16 * rb + j * 32;
        p.b_off[t * 2 + j] = (t % tb) * n * 64 + j * 32;
      }
  }
  std::vector<float> got;
  long long cyc = 0;
  if (!run(a, b, p, got, &cyc)) exit(2);
  const double per = (double)cyc / (p.repeat * p.n_mma);
  printf("T5 rate N=%3d %-18s: %8lld cycles for %d MMAs (M=128,K=16) -> %.1f cyc/MMA, %.0f%% of the 128*N/256 floor rate\n", n,
         mode == 0 ? "planar-noswizzle" : (mode == 1 ? "SW128-rows" : "SW64-rows"), cyc, p.repeat * p.n_mma, per,
         100.0 * (128.0 * n / 256.0) / per);
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no device\n"); return 1; }
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  cudaMalloc(&d_a, 256 * 1024);
  cudaMalloc(&d_b, 256 * 1024);
  cudaMalloc(&d_d, 128 * 256 * sizeof(float));
  cudaMalloc(&d_cyc, sizeof(long long));
  test_planar(32, 32);
  test_planar(64, 64);
  const int ns[] = {16, 32, 48, 64, 96, 128, 256};
  for (int n : ns) test_rate(n, 0);
  for (int n : ns) test_rate(n, 1);
  for (int n : ns) test_rate(n, 2);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int n : ns) {
      if (mode == 1 && n > 128) continue;   // 4 accumulators of n columns
      const int trips = mode == 0 ? 512 : 32;
      const int per_trip = mode == 0 ? 16 : 216;
      rate_kernel<<<1, 128, 200 * 1024>>>(n, trips, mode, d_cyc);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("rate_kernel failed\n"); return 2; }
      long long cyc = 0;
      cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
      const double per = (double)cyc / (trips * per_trip);
      printf("T%d rate N=%3d %s: %.1f cyc/MMA (M=128,K=16); 128*N/256 = %.0f; => %.0f MAC/cyc/SM\n", 6 + mode, n,
             mode == 0 ? "same-descriptor unrolled issue" : "conv-like issue loop (27 taps x 4 tiles x 2 k16)", per,
             128.0 * n / 256.0, 128.0 * n * 16.0 / per);
    }
  test_mn_major(32);
  test_mn_major(128);
  test_swizzled(32, 128, 16);
  test_swizzled(32, 128, 10);
  test_swizzled(32, 64, 16);
  test_swizzled(32, 64, 10);
  printf("probe done\n");
  return 0;
}
