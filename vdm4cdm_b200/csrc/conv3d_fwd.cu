// Conv3d ("same", stride 1, up to 27 taps) as an implicit GEMM on the 5th-gen tensor cores.
//
// Stands in for the cuDNN conv3d fprop (and, with flipped/transposed weights, dgrad) launched by
// every torch.nn.Conv3d of mltools' CUNet (networks.py:259-265 -> blocks.py:129-132, recovered in
// model_test.ipynb:684-692).
//
// GEMM view per output tile:   D[128 voxels x N] += sum_{tap, ci-chunk} A_tap[128 x KC] * W_tap[KC x N]
//   * activations are NDHWC bf16; a tile is a (TD,TH,TW) box of <= 128 voxels.  For tap (dd,dh,dw)
//     the A operand is the same box shifted by the tap offset, fetched by ONE 5-D TMA load;
//     out-of-range voxels are zero-filled by the TMA unit, which is exactly Conv3d's zero padding,
//     so no halo buffer, no im2col matrix and no boundary branches exist anywhere.
//   * weights are [tap][Cout_pad][Cin] bf16 (K-major B operand), one 3-D TMA load per k-iteration.
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N=Cout_pad, K=16) accumulates in TMEM (fp32); two
//     accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
//   * epilogue (tcgen05.ld 32x32b): + bias/conditioning row (chan_add[b][co]) + residual, bf16 or
//     fp32 store, and per-channel (sum, sumsq) GroupNorm statistics via a warp-shuffle transpose
//     reduction, accumulated per CTA in shared memory and flushed with fp64 atomics.
//   * persistent CTAs (<= one per SM) walk the tile list with a static stride.
//
// Roofline: tensor-bound; algorithmic FLOPs = 2 * taps * Cin * Cout * B*D*H*W.
#include <cudaTypedefs.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "ptx.cuh"

namespace vdm {

constexpr int kConvThreads = 192;      // 6 warps
constexpr int kMaxStages = 8;
constexpr int kTileM = 128;

struct ConvKernelParams {
  int B, D, H, W;
  int c_in, c_out, n_pad;              // n_pad = UMMA N (Cout_pad)
  int n_taps;
  int8_t tap[VDM_MAX_TAPS][3];         // (dd, dh, dw)
  int TD, TH, TW;                      // tile box
  int tiles_d, tiles_h, tiles_w;
  int n_tiles;
  int stages;                          // smem pipeline depth
  int stage_bytes, a_bytes;            // per-stage bytes, bytes of the A part
  int tmem_cols;                       // allocated TMEM columns (power of two >= 2*n_pad)
  int out_fp32;
  void* y;
  const float* chan_add;
  const int32_t* step_ptr;
  long long chan_add_step_stride;
  const __nv_bfloat16* residual;
  double* stats;
};

struct ConvShared {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  float stat_sum[4][256];
  float stat_sq[4][256];
};

// Transpose-reduce 16 per-row values over the 32 lanes of a warp: afterwards lane l (even) holds
// the column (l >> 1) total in v[0].
__device__ __forceinline__ void warp_column_sums16(float (&v)[16]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool hi = lane & 16;
    const float send = hi ? v[i] : v[i + 8];
    const float keep = hi ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(full, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 8;
    const float send = hi ? v[i] : v[i + 4];
    const float keep = hi ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(full, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 4;
    const float send = hi ? v[i] : v[i + 2];
    const float keep = hi ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(full, send, 4);
  }
  {
    const bool hi = lane & 2;
    const float send = hi ? v[0] : v[1];
    const float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(full, send, 2);
  }
  v[0] += __shfl_xor_sync(full, v[0], 1);
}

template <int KC>  // channels per k-iteration: 16 / 32 / 64  (row = 32 / 64 / 128 bytes = swizzle span)
__global__ void __launch_bounds__(kConvThreads, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ ConvKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  ConvShared* sh = reinterpret_cast<ConvShared*>(smem + (size_t)p.stages * p.stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kRowBytes = KC * 2;
  const int k_chunks = p.c_in / KC;
  const int k_iters = p.n_taps * k_chunks;
  const int rows = p.TD * p.TH * p.TW;
  const uint32_t tx_bytes = (uint32_t)(rows * kRowBytes + p.n_pad * kRowBytes);

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    ptx::prefetch_tensormap(&tmap_w);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&sh->full[s], 1);
      ptx::mbar_init(&sh->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&sh->tmem_full[a], 1);
      ptx::mbar_init(&sh->tmem_empty[a], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&sh->tmem_base, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < 4 * 256; i += 128) {
      (&sh->stat_sum[0][0])[i] = 0.f;
      (&sh->stat_sq[0][0])[i] = 0.f;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int t = tile;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int td = t % p.tiles_d;
        const int b = t / p.tiles_d;
        const int w0 = tw * p.TW, h0 = th * p.TH, d0 = td * p.TD;
        for (int tap = 0; tap < p.n_taps; ++tap) {
          const int dd = p.tap[tap][0], dh = p.tap[tap][1], dw = p.tap[tap][2];
          for (int kc = 0; kc < k_chunks; ++kc, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            ptx::mbar_wait(&sh->empty[s], ph ^ 1);
            uint8_t* a_dst = smem + (size_t)s * p.stage_bytes;
            uint8_t* b_dst = a_dst + p.a_bytes;
            ptx::mbar_arrive_expect_tx(&sh->full[s], tx_bytes);
            ptx::tma_load_5d(a_dst, &tmap_x, &sh->full[s], kc * KC, w0 + dw, h0 + dh, d0 + dd, b);
            ptx::tma_load_3d(b_dst, &tmap_w, &sh->full[s], kc * KC, 0, tap);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::make_idesc_bf16(kTileM, (uint32_t)p.n_pad);
    uint32_t it = 0;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1;
      const uint32_t aph = (ti >> 1) & 1;
      ptx::mbar_wait(&sh->tmem_empty[acc], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.n_pad;
      for (int ki = 0; ki < k_iters; ++ki, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        ptx::mbar_wait(&sh->full[s], ph);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = ptx::smem_u32(smem + (size_t)s * p.stage_bytes);
          const uint32_t b_addr = a_addr + (uint32_t)p.a_bytes;
          const uint64_t a_desc = ptx::make_kmajor_desc(a_addr, kRowBytes);
          const uint64_t b_desc = ptx::make_kmajor_desc(b_addr, kRowBytes);
#pragma unroll
          for (int j = 0; j < KC / 16; ++j) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle span: +2 in the >>4 address field
            ptx::umma_bf16(d_tmem, a_desc + (uint64_t)(2 * j), b_desc + (uint64_t)(2 * j), idesc,
                           (ki > 0 || j > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&sh->empty[s]);                       // smem slot reusable when these MMAs finish
          if (ki == k_iters - 1) ptx::umma_commit(&sh->tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (4 warps, 128 TMEM lanes) =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;                  // tile row == TMEM lane
    const int lw = r % p.TW;
    const int lh = (r / p.TW) % p.TH;
    const int ld = r / (p.TW * p.TH);
    const int n_chunks = p.n_pad >> 4;
    uint32_t ti = 0;
    int stat_b = -1;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
      int t = tile;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int td = t % p.tiles_d;
      const int b = t / p.tiles_d;
      const int w = tw * p.TW + lw, h = th * p.TH + lh, d = td * p.TD + ld;
      const bool valid = (r < rows) && (w < p.W) && (h < p.H) && (d < p.D);
      const long long vox = (((long long)b * p.D + d) * p.H + h) * p.W + w;

      if (p.stats && stat_b != b) {
        if (stat_b >= 0) {
          // flush the statistics of the previous batch sample
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int c = threadIdx.x - 64; c < p.c_out; c += 128) {
            const float s1 = sh->stat_sum[0][c] + sh->stat_sum[1][c] + sh->stat_sum[2][c] + sh->stat_sum[3][c];
            const float s2 = sh->stat_sq[0][c] + sh->stat_sq[1][c] + sh->stat_sq[2][c] + sh->stat_sq[3][c];
            atomicAdd(p.stats + ((long long)stat_b * p.c_out + c) * 2, (double)s1);
            atomicAdd(p.stats + ((long long)stat_b * p.c_out + c) * 2 + 1, (double)s2);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int i = threadIdx.x - 64; i < 4 * 256; i += 128) {
            (&sh->stat_sum[0][0])[i] = 0.f;
            (&sh->stat_sq[0][0])[i] = 0.f;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        stat_b = b;
      }

      const uint32_t acc = ti & 1;
      const uint32_t aph = (ti >> 1) & 1;
      ptx::mbar_wait(&sh->tmem_full[acc], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.n_pad;
      const float* cadd = nullptr;
      if (p.chan_add) {
        const long long step = p.step_ptr ? (long long)(*p.step_ptr) : 0ll;
        cadd = p.chan_add + step * p.chan_add_step_stride + (long long)b * p.c_out;
      }
      for (int ch = 0; ch < n_chunks; ++ch) {
        uint32_t raw[16];
        ptx::tmem_ld16(taddr + (uint32_t)(ch * 16), raw);
        ptx::tmem_ld_wait();
        const int c0 = ch * 16;
        if (c0 >= p.c_out) continue;  // padded output channels (warp-uniform)
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(raw[j]);
        const bool full16 = (c0 + 16 <= p.c_out);
        if (cadd) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (full16 || c0 + j < p.c_out) f[j] += __ldg(cadd + c0 + j);
        }
        if (p.residual && valid) {
          const __nv_bfloat16* rp = p.residual + vox * p.c_out + c0;
          if (full16) {
            float r0[8], r1[8];
            unpack8(*reinterpret_cast<const bf16x8*>(rp), r0);
            unpack8(*reinterpret_cast<const bf16x8*>(rp + 8), r1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] += r0[j];
              f[8 + j] += r1[j];
            }
          } else {
            for (int j = 0; j < 16 && c0 + j < p.c_out; ++j) f[j] += __bfloat162float(rp[j]);
          }
        }
        if (p.out_fp32) {
          if (valid) {
            float* yp = static_cast<float*>(p.y) + vox * p.c_out + c0;
            if (full16 && (p.c_out & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(yp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 16 && c0 + j < p.c_out; ++j) yp[j] = f[j];
            }
          }
        } else {
          // round to bf16 first so that the statistics describe the stored tensor
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            lo[j] = f[j];
            hi[j] = f[8 + j];
          }
          const bf16x8 plo = pack8(lo), phi = pack8(hi);
          if (valid) {
            __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(p.y) + vox * p.c_out + c0;
            if (full16 && (p.c_out & 7) == 0) {
              *reinterpret_cast<bf16x8*>(yp) = plo;
              *reinterpret_cast<bf16x8*>(yp + 8) = phi;
            } else {
              const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(&plo);
              const __nv_bfloat16* src2 = reinterpret_cast<const __nv_bfloat16*>(&phi);
              for (int j = 0; j < 16 && c0 + j < p.c_out; ++j) yp[j] = j < 8 ? src[j] : src2[j - 8];
            }
          }
          unpack8(plo, lo);
          unpack8(phi, hi);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[j] = lo[j];
            f[8 + j] = hi[j];
          }
        }
        if (p.stats) {
          float s1[16], s2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = valid ? f[j] : 0.f;
            s1[j] = x;
            s2[j] = x * x;
          }
          warp_column_sums16(s1);
          warp_column_sums16(s2);
          if ((lane & 1) == 0) {
            const int c = c0 + (lane >> 1);
            sh->stat_sum[q][c] += s1[0];
            sh->stat_sq[q][c] += s2[0];
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back
      ptx::tc_fence_before();
      ptx::mbar_arrive(&sh->tmem_empty[acc]);
    }
    if (p.stats && stat_b >= 0) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = threadIdx.x - 64; c < p.c_out; c += 128) {
        const float s1 = sh->stat_sum[0][c] + sh->stat_sum[1][c] + sh->stat_sum[2][c] + sh->stat_sum[3][c];
        const float s2 = sh->stat_sq[0][c] + sh->stat_sq[1][c] + sh->stat_sq[2][c] + sh->stat_sq[3][c];
        atomicAdd(p.stats + ((long long)stat_b * p.c_out + c) * 2, (double)s1);
        atomicAdd(p.stats + ((long long)stat_b * p.c_out + c) * 2 + 1, (double)s2);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---- host side ------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }();
  return fn;
}

struct TileShape {
  int td, th, tw;
};

// Pick the (TD,TH,TW) box of <= 128 voxels that covers the grid with the fewest tiles.
static TileShape pick_tile(int D, int H, int W) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, TileShape> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(D, H, W);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  TileShape best{1, 1, 1};
  long long best_tiles = -1;
  for (int tw = 1; tw <= W && tw <= kTileM; ++tw) {
    for (int th = 1; th <= H && tw * th <= kTileM; ++th) {
      int td = kTileM / (tw * th);
      if (td > D) td = D;
      const long long tiles = (long long)ceil_div(W, tw) * ceil_div(H, th) * ceil_div(D, td);
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && tw > best.tw)) {
        best_tiles = tiles;
        best = TileShape{td, th, tw};
      }
    }
  }
  cache[key] = best;
  return best;
}

static int num_sms() {
  static int n = []() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return kNumSMs;
    return v;
  }();
  return n;
}

template <int KC>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const ConvKernelParams& p, size_t smem_bytes,
                       int grid, cudaStream_t stream) {
  static bool configured = false;  // per template instantiation
  if (!configured) {
    VDM_CHECK_CUDA(cudaFuncSetAttribute(conv3d_igemm_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024));
    configured = true;
  }
  conv3d_igemm_kernel<KC><<<grid, kConvThreads, smem_bytes, stream>>>(tx, tw, p);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

}  // namespace vdm

using namespace vdm;

extern "C" int vdm_conv3d_fwd(const VdmConvDesc* desc, const void* x, const void* w, void* y,
                              const VdmConvEpilogue* epi, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VDM_CHECK_ARG(desc && x && w && y, "vdm_conv3d_fwd: NULL pointer argument");
  const VdmConvDesc& d = *desc;
  VDM_CHECK_ARG(d.batch >= 1 && d.depth >= 1 && d.height >= 1 && d.width >= 1, "vdm_conv3d_fwd: bad grid");
  VDM_CHECK_ARG(d.c_in >= 16 && d.c_in % 16 == 0, "vdm_conv3d_fwd: c_in=%d must be a multiple of 16", d.c_in);
  VDM_CHECK_ARG(d.c_out_pad >= 16 && d.c_out_pad % 16 == 0 && d.c_out_pad <= 256,
                "vdm_conv3d_fwd: c_out_pad=%d must be a multiple of 16 in [16,256]", d.c_out_pad);
  VDM_CHECK_ARG(d.c_out >= 1 && d.c_out <= d.c_out_pad, "vdm_conv3d_fwd: c_out=%d vs c_out_pad=%d", d.c_out,
                d.c_out_pad);
  VDM_CHECK_ARG(d.n_taps >= 1 && d.n_taps <= VDM_MAX_TAPS, "vdm_conv3d_fwd: n_taps=%d", d.n_taps);
  if (d.circular) {
    set_error("vdm_conv3d_fwd: circular padding is not implemented");
    return VDM_E_UNSUPPORTED;
  }
  if (epi && epi->residual_half) {
    set_error("vdm_conv3d_fwd: residual_half is not implemented");
    return VDM_E_UNSUPPORTED;
  }
  VDM_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "vdm_conv3d_fwd: pointers must be 16-byte aligned");
  auto encode = get_encode_fn();
  if (!encode) {
    set_error("vdm_conv3d_fwd: cuTensorMapEncodeTiled is not available from the driver");
    return VDM_E_DRIVER;
  }

  const int KC = (d.c_in % 64 == 0) ? 64 : (d.c_in % 32 == 0 ? 32 : 16);
  const CUtensorMapSwizzle swz =
      KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const TileShape ts = pick_tile(d.depth, d.height, d.width);

  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.B = d.batch; p.D = d.depth; p.H = d.height; p.W = d.width;
  p.c_in = d.c_in; p.c_out = d.c_out; p.n_pad = d.c_out_pad; p.n_taps = d.n_taps;
  for (int t = 0; t < d.n_taps; ++t)
    for (int k = 0; k < 3; ++k) {
      VDM_CHECK_ARG(d.tap_offset[t][k] >= -1 && d.tap_offset[t][k] <= 1, "vdm_conv3d_fwd: tap offset out of range");
      p.tap[t][k] = d.tap_offset[t][k];
    }
  p.TD = ts.td; p.TH = ts.th; p.TW = ts.tw;
  p.tiles_d = ceil_div(d.depth, ts.td); p.tiles_h = ceil_div(d.height, ts.th); p.tiles_w = ceil_div(d.width, ts.tw);
  const long long n_tiles = (long long)p.tiles_d * p.tiles_h * p.tiles_w * d.batch;
  VDM_CHECK_ARG(n_tiles < (1ll << 31), "vdm_conv3d_fwd: too many tiles");
  p.n_tiles = (int)n_tiles;
  p.a_bytes = kTileM * KC * 2;
  int b_bytes = d.c_out_pad * KC * 2;
  b_bytes = (b_bytes + 1023) & ~1023;
  p.stage_bytes = p.a_bytes + b_bytes;
  const int smem_budget = 227 * 1024 - 1024 /*alignment slack*/ - (int)sizeof(ConvShared);
  int stages = smem_budget / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  VDM_CHECK_ARG(stages >= 2, "vdm_conv3d_fwd: tile does not fit in shared memory");
  p.stages = stages;
  int cols = 32;
  while (cols < 2 * d.c_out_pad) cols <<= 1;
  p.tmem_cols = cols;
  p.out_fp32 = d.out_fp32;
  p.y = y;
  if (epi) {
    p.chan_add = epi->chan_add;
    p.step_ptr = epi->step_ptr;
    p.chan_add_step_stride = epi->chan_add_step_stride;
    p.residual = static_cast<const __nv_bfloat16*>(epi->residual);
    p.stats = epi->stats;
  }
  VDM_CHECK_ARG(!(p.stats && d.out_fp32), "vdm_conv3d_fwd: stats are only produced for bf16 outputs");

  // activations: 5-D (C, W, H, D, B), box (KC, TW, TH, TD, 1); out-of-bounds -> zeros
  CUtensorMap tmx, tmw;
  {
    cuuint64_t gdim[5] = {(cuuint64_t)d.c_in, (cuuint64_t)d.width, (cuuint64_t)d.height, (cuuint64_t)d.depth,
                          (cuuint64_t)d.batch};
    cuuint64_t gstr[4] = {(cuuint64_t)d.c_in * 2, (cuuint64_t)d.width * d.c_in * 2,
                          (cuuint64_t)d.height * d.width * d.c_in * 2,
                          (cuuint64_t)d.depth * d.height * d.width * d.c_in * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)ts.tw, (cuuint32_t)ts.th, (cuuint32_t)ts.td, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("vdm_conv3d_fwd: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
      return VDM_E_DRIVER;
    }
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)d.c_in, (cuuint64_t)d.c_out_pad, (cuuint64_t)d.n_taps};
    cuuint64_t gstr[2] = {(cuuint64_t)d.c_in * 2, (cuuint64_t)d.c_out_pad * d.c_in * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)d.c_out_pad, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("vdm_conv3d_fwd: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
      return VDM_E_DRIVER;
    }
  }

  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + sizeof(ConvShared) + 1024;
  int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  switch (KC) {
    case 64: return launch_conv<64>(tmx, tmw, p, smem_bytes, grid, stream);
    case 32: return launch_conv<32>(tmx, tmw, p, smem_bytes, grid, stream);
    default: return launch_conv<16>(tmx, tmw, p, smem_bytes, grid, stream);
  }
}
