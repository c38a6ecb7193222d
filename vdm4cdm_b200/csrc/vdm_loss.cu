// The continuous-time VDM training loss around the denoiser call, as five bandwidth-bound kernels with hand-derived
// gradients.
//
// Stands in for the ~40 eager ATen launches of VDM.get_loss and its autograd backward (mltools vdm_model.py:429-442 in
// model_test.ipynb:680; Kingma et al. 2021, eqs. 11-17, restated in oracle/vdm_ref.py:158-184):
//     gamma(t) = b + |w| t (learned_linear) or b + w t (fixed_linear)
//     z_t      = alpha_t x + sigma_t eps,  alpha^2 = sigmoid(-gamma_t), sigma^2 = sigmoid(gamma_t)        (vdm_loss_zt)
//     diffusion_b = 1/2 gamma' S_d,        S_d  = sum (eps - eps_hat)^2                                    (vdm_loss_sums)
//     latent_b    = 1/2 (N v1 + (1 - v1) S_x - N log v1 - N),   v1 = sigmoid(gamma(1)), S_x = sum x^2
//     recons_b    = 1/2 e^{gamma(0)} S_n0 / dn^2 + N (log dn + 1/2 log 2 pi),  S_n0 = sum noise0^2  (x - z_0 = -e^{gamma_0 / 2} noise0)
//     loss        = mean_b (diffusion + latent + recons) / (N log 2)                                       (vdm_loss_finalize)
// and backward
//     d eps_hat = g k gamma' (eps_hat - eps),  k = 1 / (B N log 2)                                         (vdm_loss_dpred)
//     d b, d w from the three sums (finalize) and, through z_t, from S1 = sum g_zt x, S2 = sum g_zt eps:
//     d gamma_t = -1/2 alpha sigma^2 S1 + 1/2 sigma alpha^2 S2                                             (vdm_loss_zt_bwd)
// HBM traffic: 12 B/voxel (zt), 16 (sums), 12 (dpred), 12 (zt_bwd): the algorithmic minimum of each pass.
#include "common.cuh"

namespace vdm {

__device__ __forceinline__ double gamma_of(const float* gb, const float* gw, int learned, double t) {
  const double w = learned ? fabs((double)*gw) : (double)*gw;
  return (double)*gb + w * t;
}
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// block-wide sum of NV doubles per thread -> atomics into out[0..NV)
template <int NV>
__device__ __forceinline__ void block_atomic_sums(double (&v)[NV], double* __restrict__ out) {
  __shared__ double s_part[8][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) s_part[threadIdx.x >> 5][i] = v[i];
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}

// grid (blocks, B): z_t = alpha_b x + sigma_b eps
__global__ void __launch_bounds__(256)
vdm_zt_kernel(const float4* __restrict__ x, const float4* __restrict__ eps, const float* __restrict__ times, const float* gb,
              const float* gw, int learned, float4* __restrict__ zt, int64_t n4) {
  const int b = blockIdx.y;
  const double g = gamma_of(gb, gw, learned, (double)times[b]);
  const float alpha = (float)sqrt(sigmoid_d(-g)), sigma = (float)sqrt(sigmoid_d(g));
  const float4* xb = x + (int64_t)b * n4;
  const float4* eb = eps + (int64_t)b * n4;
  float4* zb = zt + (int64_t)b * n4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 a = __ldg(xb + i), e = __ldg(eb + i);
    zb[i] = make_float4(fmaf(alpha, a.x, sigma * e.x), fmaf(alpha, a.y, sigma * e.y), fmaf(alpha, a.z, sigma * e.z),
                        fmaf(alpha, a.w, sigma * e.w));
  }
}

// grid (blocks, B): sums[b] += (sum g x, sum g eps)
__global__ void __launch_bounds__(256)
vdm_zt_bwd_kernel(const float4* __restrict__ g, const float4* __restrict__ x, const float4* __restrict__ eps,
                  double* __restrict__ sums, int64_t n4) {
  const int b = blockIdx.y;
  const float4* gp = g + (int64_t)b * n4;
  const float4* xb = x + (int64_t)b * n4;
  const float4* eb = eps + (int64_t)b * n4;
  double v[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 gg = __ldg(gp + i), a = __ldg(xb + i), e = __ldg(eb + i);
    v[0] += (double)(gg.x * a.x + gg.y * a.y) + (double)(gg.z * a.z + gg.w * a.w);
    v[1] += (double)(gg.x * e.x + gg.y * e.y) + (double)(gg.z * e.z + gg.w * e.w);
  }
  block_atomic_sums<2>(v, sums + 2 * b);
}

// one block: d b, d w of the z_t path from the per-sample sums
__global__ void vdm_zt_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ times, const float* gb,
                                           const float* gw, int learned, int B, float* __restrict__ out2) {
  if (threadIdx.x != 0) return;
  double db = 0.0, dw = 0.0;
  for (int b = 0; b < B; ++b) {
    const double t = (double)times[b];
    const double g = gamma_of(gb, gw, learned, t);
    const double s2 = sigmoid_d(g), a2 = 1.0 - s2;                   // sigma^2, alpha^2
    const double alpha = sqrt(a2), sigma = sqrt(s2);
    const double dg = -0.5 * alpha * s2 * sums[2 * b] + 0.5 * sigma * a2 * sums[2 * b + 1];
    db += dg;
    dw += dg * t;
  }
  if (learned) dw *= (*gw < 0.f ? -1.0 : (*gw > 0.f ? 1.0 : 0.0));   // d|w|/dw (torch.abs: 0 at 0)
  out2[0] = (float)db;
  out2[1] = (float)dw;
}

// grid (blocks, B): sums[b] += (sum (eps - pred)^2, sum x^2, sum noise0^2)
__global__ void __launch_bounds__(256)
vdm_loss_sums_kernel(const float4* __restrict__ pred, const float4* __restrict__ eps, const float4* __restrict__ x,
                     const float4* __restrict__ n0, double* __restrict__ sums, int64_t n4) {
  const int b = blockIdx.y;
  const int64_t off = (int64_t)b * n4;
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 p = __ldg(pred + off + i), e = __ldg(eps + off + i), a = __ldg(x + off + i), z = __ldg(n0 + off + i);
    const float d0 = e.x - p.x, d1 = e.y - p.y, d2 = e.z - p.z, d3 = e.w - p.w;
    v[0] += (double)(d0 * d0 + d1 * d1) + (double)(d2 * d2 + d3 * d3);
    v[1] += (double)(a.x * a.x + a.y * a.y) + (double)(a.z * a.z + a.w * a.w);
    v[2] += (double)(z.x * z.x + z.y * z.y) + (double)(z.z * z.z + z.w * z.w);
  }
  block_atomic_sums<3>(v, sums + 3 * b);
}

// one block: loss and its three terms (bits per dimension, batch means), the scalar of d eps_hat, and d b, d w of the loss terms
// out8 = (loss, diffusion, latent, recons, k gamma', d b, d w, unused)
__global__ void vdm_loss_finalize_kernel(const double* __restrict__ sums, const float* gb, const float* gw, int learned, int B,
                                         double n_vox, double data_noise, float* __restrict__ out8) {
  if (threadIdx.x != 0) return;
  const double wabs = learned ? fabs((double)*gw) : (double)*gw;     // gamma'
  const double g0 = (double)*gb, g1 = g0 + wabs;
  const double v1 = sigmoid_d(g1), v1p = v1 * (1.0 - v1);
  const double e0 = exp(g0) / (data_noise * data_noise);
  const double k = 1.0 / ((double)B * n_vox * log(2.0));
  double diff = 0.0, lat = 0.0, rec = 0.0, dlat = 0.0, drec = 0.0, sd = 0.0;
  for (int b = 0; b < B; ++b) {
    const double S_d = sums[3 * b], S_x = sums[3 * b + 1], S_n = sums[3 * b + 2];
    diff += 0.5 * wabs * S_d;
    sd += 0.5 * S_d;
    lat += 0.5 * (n_vox * v1 + (1.0 - v1) * S_x - n_vox * log(v1) - n_vox);
    dlat += 0.5 * (n_vox * v1p - v1p * S_x - n_vox * v1p / v1);      // d latent / d gamma_1
    rec += 0.5 * e0 * S_n + n_vox * (log(data_noise) + 0.5 * log(2.0 * 3.14159265358979323846));
    drec += 0.5 * e0 * S_n;                                          // d recons / d gamma_0
  }
  out8[0] = (float)((diff + lat + rec) * k);
  out8[1] = (float)(diff * k);
  out8[2] = (float)(lat * k);
  out8[3] = (float)(rec * k);
  out8[4] = (float)(k * wabs);
  double db = (dlat + drec) * k, dw = (sd + dlat) * k;               // gamma_1 = b + |w|, gamma_0 = b, gamma' = |w|
  if (learned) dw *= (*gw < 0.f ? -1.0 : (*gw > 0.f ? 1.0 : 0.0));
  out8[5] = (float)db;
  out8[6] = (float)dw;
  out8[7] = 0.f;
}

// d eps_hat = g * coef * (eps_hat - eps)
__global__ void __launch_bounds__(256)
vdm_dpred_kernel(const float4* __restrict__ pred, const float4* __restrict__ eps, const float* __restrict__ coef,
                 const float* __restrict__ g_loss, float4* __restrict__ d_pred, int64_t n4) {
  const float c = *coef * *g_loss;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 p = __ldg(pred + i), e = __ldg(eps + i);
    d_pred[i] = make_float4(c * (p.x - e.x), c * (p.y - e.y), c * (p.z - e.z), c * (p.w - e.w));
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline dim3 loss_grid(int64_t n4, int batch) {
  int64_t blocks = (n4 + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = ((int64_t)kNumSMs * 8 + batch - 1) / batch;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks, (unsigned)batch);
}

}  // namespace vdm

using namespace vdm;

extern "C" int vdm_loss_zt(const float* x, const float* noise, const float* times, const float* gamma_b, const float* gamma_w,
                           int learned, float* zt, int batch, int64_t n, void* stream) {
  VDM_CHECK_ARG(x && noise && times && gamma_b && gamma_w && zt && batch >= 1 && batch <= 65535 && n >= 4 && n % 4 == 0,
                "vdm_loss_zt: bad argument (n must be a multiple of 4)");
  VDM_CHECK_ARG(aligned16(x) && aligned16(noise) && aligned16(zt), "vdm_loss_zt: pointers must be 16-byte aligned");
  vdm_zt_kernel<<<loss_grid(n / 4, batch), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(noise), times, gamma_b, gamma_w, learned,
      reinterpret_cast<float4*>(zt), n / 4);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

extern "C" int vdm_loss_zt_bwd(const float* g_zt, const float* x, const float* noise, const float* times, const float* gamma_b,
                               const float* gamma_w, int learned, double* work, float* grads2, int batch, int64_t n, void* stream) {
  VDM_CHECK_ARG(g_zt && x && noise && times && gamma_b && gamma_w && work && grads2 && batch >= 1 && batch <= 65535 && n >= 4 &&
                    n % 4 == 0, "vdm_loss_zt_bwd: bad argument");
  VDM_CHECK_ARG(aligned16(g_zt) && aligned16(x) && aligned16(noise), "vdm_loss_zt_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  VDM_CHECK_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * 2 * batch, st));
  vdm_zt_bwd_kernel<<<loss_grid(n / 4, batch), 256, 0, st>>>(reinterpret_cast<const float4*>(g_zt), reinterpret_cast<const float4*>(x),
                                                             reinterpret_cast<const float4*>(noise), work, n / 4);
  VDM_CHECK_LAUNCH();
  vdm_zt_bwd_finalize_kernel<<<1, 32, 0, st>>>(work, times, gamma_b, gamma_w, learned, batch, grads2);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

extern "C" int vdm_loss_terms(const float* pred, const float* noise, const float* x, const float* noise0, const float* gamma_b,
                              const float* gamma_w, int learned, double data_noise, double* work, float* out8, int batch, int64_t n,
                              void* stream) {
  VDM_CHECK_ARG(pred && noise && x && noise0 && gamma_b && gamma_w && work && out8 && batch >= 1 && batch <= 65535 && n >= 4 &&
                    n % 4 == 0 && data_noise > 0.0, "vdm_loss_terms: bad argument");
  VDM_CHECK_ARG(aligned16(pred) && aligned16(noise) && aligned16(x) && aligned16(noise0), "vdm_loss_terms: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  VDM_CHECK_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * 3 * batch, st));
  vdm_loss_sums_kernel<<<loss_grid(n / 4, batch), 256, 0, st>>>(
      reinterpret_cast<const float4*>(pred), reinterpret_cast<const float4*>(noise), reinterpret_cast<const float4*>(x),
      reinterpret_cast<const float4*>(noise0), work, n / 4);
  VDM_CHECK_LAUNCH();
  vdm_loss_finalize_kernel<<<1, 32, 0, st>>>(work, gamma_b, gamma_w, learned, batch, (double)n, data_noise, out8);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

extern "C" int vdm_loss_dpred(const float* pred, const float* noise, const float* coef, const float* g_loss, float* d_pred,
                              int64_t total, void* stream) {
  VDM_CHECK_ARG(pred && noise && coef && g_loss && d_pred && total >= 4 && total % 4 == 0, "vdm_loss_dpred: bad argument");
  VDM_CHECK_ARG(aligned16(pred) && aligned16(noise) && aligned16(d_pred), "vdm_loss_dpred: pointers must be 16-byte aligned");
  int64_t blocks = (total / 4 + 256 * 4 - 1) / (256 * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  vdm_dpred_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(pred),
                                                                       reinterpret_cast<const float4*>(noise), coef, g_loss,
                                                                       reinterpret_cast<float4*>(d_pred), total / 4);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}
