// Fused elementwise passes of the ResNet blocks (all HBM-bound, NDHWC bf16, 16-byte accesses).
//
// Stand in for the ATen group_norm / silu / dropout / avg_pool3d / interpolate(nearest) / cat
// launches issued by mltools' ResNetBlock / ResNetDown (blocks.py:129-170 in the traceback of
// model_test.ipynb:686-692).  The reference runs each as its own full-tensor pass; here
//   * GroupNorm statistics are per-channel (sum, sumsq) pairs produced by whichever kernel
//     writes the tensor (conv epilogue, pool, concat) and reduced to groups on the fly,
//   * normalise + affine + SiLU (+ dropout) is one read and one write,
//   * up-sampling and concat are one write.
//
// Thread mapping used everywhere: the tensor is a flat array of 16-byte chunks (8 channels);
// blockDim = 384 is a multiple of every chunks-per-voxel count in use (2,4,8,12,16,24,32,48), so a
// thread keeps the same 8 channels for its whole grid-stride loop and can hold their
// scale/shift or partial sums in registers.
#include "common.cuh"

namespace vdm {

constexpr int kEwThreads = 384;

// Reduce per-thread (sum[8], sq[8]) over the threads of a block that share a channel chunk and
// add the result to stats[c][2] (double) with one atomic per channel and quantity.
__device__ __forceinline__ void block_flush_stats(const float (&sum)[8], const float (&sq)[8], int cg_count,
                                                  double* __restrict__ stats_b) {
  __shared__ float s_part[16][kEwThreads + 1];
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_part[j][tid] = sum[j];
    s_part[8 + j][tid] = sq[j];
  }
  __syncthreads();
  // thread t handles (quantity q, channel c): q = t / C, c = t % C, C = 8*cg_count
  const int C = 8 * cg_count;
  for (int t = tid; t < 2 * C; t += kEwThreads) {
    const int q = t / C, c = t % C;
    const int cg = c >> 3, j = c & 7;
    float acc = 0.f;
    for (int p = cg; p < kEwThreads; p += cg_count) acc += s_part[q * 8 + j][p];
    atomicAdd(stats_b + 2 * c + q, (double)acc);
  }
  __syncthreads();
}

// ---- channel statistics -------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
channel_stats_kernel(const bf16x8* __restrict__ x, int64_t chunks_per_sample, int cg_count,
                     double* __restrict__ stats) {
  const int b = blockIdx.y;
  const bf16x8* xb = x + (int64_t)b * chunks_per_sample;
  float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < chunks_per_sample;
       i += (int64_t)gridDim.x * kEwThreads) {
    float f[8];
    unpack8(xb[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sum[j] += f[j];
      sq[j] += f[j] * f[j];
    }
  }
  block_flush_stats(sum, sq, cg_count, stats + (int64_t)b * cg_count * 16);
}

// ---- GroupNorm scale/shift from channel stats ------------------------------------------------
// s_scale/s_shift: shared float[C].  y = x*scale + shift  ==  gamma*(x-mean)*rstd + beta.
__device__ __forceinline__ void group_scale_shift(const double* __restrict__ stats_b, int C, int groups,
                                                  double count_per_channel, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, float eps, float* s_scale,
                                                  float* s_shift) {
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / cpg) * cpg;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < cpg; ++k) {
      s1 += stats_b[2 * (g0 + k)];
      s2 += stats_b[2 * (g0 + k) + 1];
    }
    const double n = count_per_channel * cpg;
    const double mean = s1 / n;
    double var = s2 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * rstd;
    s_scale[c] = sc;
    s_shift[c] = beta[c] - (float)mean * sc;
  }
}

// 8 keep-flags for chunk `chunk` of dropout layer `tag`: bit j set = keep element j.
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t chunk, uint32_t tag, uint64_t seed, uint32_t thresh16) {
  const uint4 w = philox4x32_10(make_uint4((uint32_t)chunk, (uint32_t)(chunk >> 32), tag, kStreamTagDropout),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t r[4] = {w.x, w.y, w.z, w.w};
  uint32_t keep = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep |= ((r[i] & 0xffffu) >= thresh16 ? 1u : 0u) << (2 * i);
    keep |= ((r[i] >> 16) >= thresh16 ? 1u : 0u) << (2 * i + 1);
  }
  return keep;
}

__global__ void __launch_bounds__(kEwThreads)
gn_silu_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ y, int64_t voxels, int C, int groups,
               const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
               float eps, float dropout_p, uint64_t seed, uint32_t layer_tag) {
  extern __shared__ float s_ss[];  // scale[C], shift[C]
  float* s_scale = s_ss;
  float* s_shift = s_ss + C;
  const int b = blockIdx.y;
  const int cg_count = C >> 3;
  group_scale_shift(stats + (int64_t)b * C * 2, C, groups, (double)voxels, gamma, beta, eps, s_scale, s_shift);
  __syncthreads();
  const int cg = threadIdx.x % cg_count;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_scale[cg * 8 + j];
    sh[j] = s_shift[cg * 8 + j];
  }
  const int64_t chunks = voxels * cg_count;
  const bf16x8* xb = x + (int64_t)b * chunks;
  bf16x8* yb = y + (int64_t)b * chunks;
  const bool drop = dropout_p > 0.f;
  const uint32_t thresh16 = (uint32_t)(dropout_p * 65536.0f + 0.5f);
  const float keep_scale = drop ? 1.0f / (1.0f - dropout_p) : 1.0f;
  for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < chunks;
       i += (int64_t)gridDim.x * kEwThreads) {
    float f[8];
    unpack8(xb[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = silu(f[j] * sc[j] + sh[j]);
    if (drop) {
      const uint32_t keep = dropout_keep8((uint64_t)b * chunks + i, layer_tag, seed, thresh16);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = ((keep >> j) & 1u) ? f[j] * keep_scale : 0.f;
    }
    yb[i] = pack8(f);
  }
}

// ---- 2x2x2 average pooling (+ stats of the output) ---------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
avgpool2_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ y, int D, int H, int W, int cg_count,
                double* __restrict__ stats) {
  const int b = blockIdx.y;
  const int Do = D >> 1, Ho = H >> 1, Wo = W >> 1;
  const int64_t out_chunks = (int64_t)Do * Ho * Wo * cg_count;
  const bf16x8* xb = x + (int64_t)b * D * H * W * cg_count;
  bf16x8* yb = y + (int64_t)b * out_chunks;
  float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < out_chunks;
       i += (int64_t)gridDim.x * kEwThreads) {
    const int cg = (int)(i % cg_count);
    int64_t v = i / cg_count;
    const int wo = (int)(v % Wo); v /= Wo;
    const int ho = (int)(v % Ho);
    const int dz = (int)(v / Ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int d = 2 * dz + (k >> 2), h = 2 * ho + ((k >> 1) & 1), w = 2 * wo + (k & 1);
      float f[8];
      unpack8(xb[(((int64_t)d * H + h) * W + w) * cg_count + cg], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.125f;
    const bf16x8 packed = pack8(acc);
    yb[i] = packed;
    if (stats) {
      float r[8];
      unpack8(packed, r);  // statistics of the stored (rounded) tensor
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum[j] += r[j];
        sq[j] += r[j] * r[j];
      }
    }
  }
  if (stats) block_flush_stats(sum, sq, cg_count, stats + (int64_t)b * cg_count * 16);
}

// ---- nearest x2 up-sampling + channel concat (+ stats of the output) ---------------------------
__global__ void __launch_bounds__(kEwThreads)
upsample_concat_kernel(const bf16x8* __restrict__ coarse, const bf16x8* __restrict__ skip, bf16x8* __restrict__ y,
                       int D, int H, int W, int cg_coarse, int cg_skip, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const int cg_count = cg_coarse + cg_skip;
  const int Dc = D >> 1, Hc = H >> 1, Wc = W >> 1;
  const int64_t voxels = (int64_t)D * H * W;
  const int64_t chunks = voxels * cg_count;
  const bf16x8* cb = coarse + (int64_t)b * Dc * Hc * Wc * cg_coarse;
  const bf16x8* sb = skip + (int64_t)b * voxels * cg_skip;
  bf16x8* yb = y + (int64_t)b * chunks;
  float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < chunks;
       i += (int64_t)gridDim.x * kEwThreads) {
    const int cg = (int)(i % cg_count);
    const int64_t v = i / cg_count;
    bf16x8 val;
    if (cg < cg_coarse) {
      int64_t t = v;
      const int w = (int)(t % W); t /= W;
      const int h = (int)(t % H);
      const int d = (int)(t / H);
      val = cb[(((int64_t)(d >> 1) * Hc + (h >> 1)) * Wc + (w >> 1)) * cg_coarse + cg];
    } else {
      val = sb[v * cg_skip + (cg - cg_coarse)];
    }
    yb[i] = val;
    if (stats) {
      float r[8];
      unpack8(val, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum[j] += r[j];
        sq[j] += r[j] * r[j];
      }
    }
  }
  if (stats) block_flush_stats(sum, sq, cg_count, stats + (int64_t)b * cg_count * 16);
}

static inline bool chunk_count_ok(int channels) {
  return channels >= 8 && channels % 8 == 0 && kEwThreads % (channels / 8) == 0;
}

static inline dim3 ew_grid(int64_t chunks, int batch) {
  int64_t blocks = (chunks + kEwThreads * 4 - 1) / (kEwThreads * 4);
  const int64_t cap = ((int64_t)kNumSMs * 5 + batch - 1) / batch;  // 5 CTAs of 384 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks, (unsigned)batch);
}

}  // namespace vdm

using namespace vdm;

extern "C" int vdm_channel_stats(const void* x, int batch, int64_t voxels, int channels, double* stats,
                                 void* stream) {
  VDM_CHECK_ARG(x && stats && batch >= 1 && batch <= 65535 && voxels >= 1, "vdm_channel_stats: bad argument");
  VDM_CHECK_ARG(chunk_count_ok(channels), "vdm_channel_stats: unsupported channel count %d", channels);
  const int cg = channels / 8;
  channel_stats_kernel<<<ew_grid(voxels * cg, batch), kEwThreads, 0, (cudaStream_t)stream>>>(
      static_cast<const bf16x8*>(x), voxels * cg, cg, stats);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

extern "C" int vdm_gn_silu(const void* x, void* y, int batch, int64_t voxels, int channels, int groups,
                           const double* stats, const float* gamma, const float* beta, float eps, float dropout_p,
                           uint64_t seed, uint32_t layer_tag, void* stream) {
  VDM_CHECK_ARG(x && y && stats && gamma && beta, "vdm_gn_silu: NULL pointer argument");
  VDM_CHECK_ARG(batch >= 1 && batch <= 65535 && voxels >= 1, "vdm_gn_silu: bad shape");
  VDM_CHECK_ARG(chunk_count_ok(channels), "vdm_gn_silu: unsupported channel count %d", channels);
  VDM_CHECK_ARG(groups >= 1 && channels % groups == 0, "vdm_gn_silu: %d channels not divisible into %d groups",
                channels, groups);
  VDM_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "vdm_gn_silu: dropout_p %f out of [0,1)", dropout_p);
  const int cg = channels / 8;
  gn_silu_kernel<<<ew_grid(voxels * cg, batch), kEwThreads, 2 * channels * sizeof(float), (cudaStream_t)stream>>>(
      static_cast<const bf16x8*>(x), static_cast<bf16x8*>(y), voxels, channels, groups, stats, gamma, beta, eps,
      dropout_p, seed, layer_tag);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

extern "C" int vdm_avgpool2(const void* x, void* y, int batch, int depth, int height, int width, int channels,
                            double* stats, void* stream) {
  VDM_CHECK_ARG(x && y && batch >= 1 && batch <= 65535, "vdm_avgpool2: bad argument");
  VDM_CHECK_ARG(depth >= 2 && height >= 2 && width >= 2 && depth % 2 == 0 && height % 2 == 0 && width % 2 == 0,
                "vdm_avgpool2: grid (%d,%d,%d) must be even", depth, height, width);
  VDM_CHECK_ARG(chunk_count_ok(channels), "vdm_avgpool2: unsupported channel count %d", channels);
  const int cg = channels / 8;
  const int64_t out_chunks = (int64_t)(depth / 2) * (height / 2) * (width / 2) * cg;
  avgpool2_kernel<<<ew_grid(out_chunks, batch), kEwThreads, 0, (cudaStream_t)stream>>>(
      static_cast<const bf16x8*>(x), static_cast<bf16x8*>(y), depth, height, width, cg, stats);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

extern "C" int vdm_upsample_concat(const void* coarse, const void* skip, void* y, int batch, int depth, int height,
                                   int width, int c_coarse, int c_skip, double* stats, void* stream) {
  VDM_CHECK_ARG(coarse && skip && y && batch >= 1 && batch <= 65535, "vdm_upsample_concat: bad argument");
  VDM_CHECK_ARG(depth % 2 == 0 && height % 2 == 0 && width % 2 == 0 && depth >= 2 && height >= 2 && width >= 2,
                "vdm_upsample_concat: fine grid (%d,%d,%d) must be even", depth, height, width);
  VDM_CHECK_ARG(c_coarse % 8 == 0 && c_skip % 8 == 0 && chunk_count_ok(c_coarse + c_skip),
                "vdm_upsample_concat: unsupported channel counts %d + %d", c_coarse, c_skip);
  const int64_t chunks = (int64_t)depth * height * width * ((c_coarse + c_skip) / 8);
  upsample_concat_kernel<<<ew_grid(chunks, batch), kEwThreads, 0, (cudaStream_t)stream>>>(
      static_cast<const bf16x8*>(coarse), static_cast<const bf16x8*>(skip), static_cast<bf16x8*>(y), depth, height,
      width, c_coarse / 8, c_skip / 8, stats);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}
