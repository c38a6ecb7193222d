// P(k) / r(k): cuFFT R2C followed by a single-pass k-shell binning kernel.
//
// Stands in for power() at src/utils.py:16-83 of the reference (pk :85-102, get_ccs :110-128):
// rfftn -> X conj(X2) -> batch mean / channel sum -> ceil(|k|) bins with Hermitian weights.
// The reference builds ~10 full-size temporaries and calls torch.bincount three times; here the
// wave number is recomputed from the flat mode index, equal-bin runs inside a warp are combined
// with a segmented shuffle scan (warp-aggregated atomics), blocks accumulate in shared memory
// and flush once with fp64 global atomics.  k-mean and mode counts depend on the grid only and
// are produced by a separate read-free geometry kernel.
//
// HBM roofline: the binning kernel reads each complex mode once: 8 (auto) or 16 (cross) bytes
// per mode per transform, N0*N1*(N2/2+1) modes.
#include <cufft.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace vdm {

struct PkGeom {
  int n0, n1, n2, n2h;  // grid; n2h = n2/2+1 modes on the half axis
  int kmax;             // bins 1..kmax are kept
  int last_even;        // Nyquist plane of the half axis has weight 1
  int64_t modes;        // n0*n1*n2h
};

__device__ __forceinline__ void mode_bin(const PkGeom& g, int64_t m, int& bin, int& weight, float& kmag) {
  const int i2 = (int)(m % g.n2h);
  const int64_t r = m / g.n2h;
  const int i1 = (int)(r % g.n1);
  const int i0 = (int)(r / g.n1);
  const int f0 = i0 > g.n0 / 2 ? i0 - g.n0 : i0;
  const int f1 = i1 > g.n1 / 2 ? i1 - g.n1 : i1;
  const int k2 = f0 * f0 + f1 * f1 + i2 * i2;
  kmag = sqrtf((float)k2);
  bin = (int)ceilf(kmag);
  weight = (i2 == 0 || (g.last_even && i2 == g.n2h - 1)) ? 1 : 2;
}

// Segmented inclusive scan over runs of equal `bin` in a warp; returns true on the tail lane of a run.
template <int NV>
__device__ __forceinline__ bool warp_run_sum(int bin, float (&v)[NV]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(full, bin, 1);
  int head = (lane == 0) || (prev != bin);
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float up[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) up[i] = __shfl_up_sync(full, v[i], o);
    const int head_up = __shfl_up_sync(full, head, o);
    if (lane >= o && !head) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += up[i];
      head |= head_up;
    }
  }
  const int next = __shfl_down_sync(full, bin, 1);
  return (lane == 31) || (next != bin);
}

// acc: double [n_fields][NSPEC][kmax+1], zero-initialised.
// NSPEC == 1: Re(X conj(X2)) (X2 may alias X).  NSPEC == 3: |X|^2, |X2|^2, Re(X conj(X2)).
template <int NSPEC>
__global__ void __launch_bounds__(256)
pk_bin_kernel(const float2* __restrict__ X, const float2* __restrict__ X2, PkGeom g, int n_transforms,
              double* __restrict__ acc) {
  extern __shared__ double s_acc[];  // [warp][NSPEC][kmax+1]: private bins per warp
  const int nb = g.kmax + 1;
  const int n_warps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < n_warps * NSPEC * nb; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  double* my = s_acc + (threadIdx.x >> 5) * NSPEC * nb;

  // One warp per (i0, i1) row of the half spectrum: the row's n2h modes are contiguous, (f0, f1) are computed once
  // per row instead of two 64-bit divisions per mode, and |k| grows along the row, so equal-bin runs are long.
  // Four 32-mode trips are loaded before any is reduced (four 8-byte loads in flight per lane).
  // (r02l: the one-mode-per-thread version ran at 0.9 TB/s, instruction-bound.)
  const int field = blockIdx.y;
  const float2* x = X + (int64_t)field * n_transforms * g.modes;
  const float2* x2 = X2 + (int64_t)field * n_transforms * g.modes;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int rows = g.n0 * g.n1;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    const int i1 = row % g.n1, i0 = row / g.n1;
    const int f0 = i0 > g.n0 / 2 ? i0 - g.n0 : i0;
    const int f1 = i1 > g.n1 / 2 ? i1 - g.n1 : i1;
    const int k01 = f0 * f0 + f1 * f1;
    const float2* xr = x + (int64_t)row * g.n2h;
    const float2* xr2 = x2 + (int64_t)row * g.n2h;
    for (int base = 0; base < g.n2h; base += 128) {
      int bins[4];
      float v[4][NSPEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i2 = base + u * 32 + lane;
        bins[u] = -1;
#pragma unroll
        for (int i = 0; i < NSPEC; ++i) v[u][i] = 0.f;
        if (i2 < g.n2h) {
          const int b = (int)ceilf(sqrtf((float)(k01 + i2 * i2)));
          if (b >= 1 && b <= g.kmax) {
            bins[u] = b;
            for (int t = 0; t < n_transforms; ++t) {
              const float2 a = __ldg(xr + (int64_t)t * g.modes + i2);
              const float2 c = __ldg(xr2 + (int64_t)t * g.modes + i2);
              if constexpr (NSPEC == 1) {
                v[u][0] += a.x * c.x + a.y * c.y;
              } else {
                v[u][0] += a.x * a.x + a.y * a.y;
                v[u][NSPEC - 2] += c.x * c.x + c.y * c.y;
                v[u][NSPEC - 1] += a.x * c.x + a.y * c.y;
              }
            }
            const float w = (i2 == 0 || (g.last_even && i2 == g.n2h - 1)) ? 1.f : 2.f;
#pragma unroll
            for (int i = 0; i < NSPEC; ++i) v[u][i] *= w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (base + u * 32 >= g.n2h) break;           // warp-uniform
        // |k| grows along a row, so after the run merge every bin has at most ONE tail lane in this trip: the
        // warp's private bins take plain read-modify-writes (fp64 shared-memory atomics are compare-and-swap loops;
        // with eight warps of neighbouring rows hitting the same few bins they were the cost of this kernel, r02m)
        const bool tail = warp_run_sum<NSPEC>(bins[u], v[u]);
        if (tail && bins[u] >= 1) {
#pragma unroll
          for (int i = 0; i < NSPEC; ++i) my[i * nb + bins[u]] += (double)v[u][i];
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  double* out = acc + (int64_t)field * NSPEC * nb;
  for (int i = threadIdx.x; i < NSPEC * nb; i += blockDim.x) {
    double s = 0.0;
    for (int wv = 0; wv < n_warps; ++wv) s += s_acc[wv * NSPEC * nb + i];
    if (s != 0.0) atomicAdd(out + i, s);
  }
}

// geometry: ksum[bin] += w*|k| (double), cnt[bin] += w (unsigned long long). No memory reads.
__global__ void __launch_bounds__(256)
pk_geometry_kernel(PkGeom g, double* __restrict__ ksum, unsigned long long* __restrict__ cnt) {
  extern __shared__ double s_k[];  // [kmax+1] doubles then [kmax+1] uint64
  const int nb = g.kmax + 1;
  unsigned long long* s_c = reinterpret_cast<unsigned long long*>(s_k + nb);
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    s_k[i] = 0.0;
    s_c[i] = 0ull;
  }
  __syncthreads();
  const int64_t span = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (g.modes + span - 1) / span;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t m = it * span + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bin = -1, w = 0;
    float kmag = 0.f;
    float v[2] = {0.f, 0.f};
    if (m < g.modes) {
      mode_bin(g, m, bin, w, kmag);
      if (bin >= 1 && bin <= g.kmax) {
        v[0] = kmag * (float)w;
        v[1] = (float)w;  // <= 64 per run: exact in fp32
      } else {
        bin = -1;
      }
    }
    const bool tail = warp_run_sum<2>(bin, v);
    if (tail && bin >= 1) {
      atomicAdd(&s_k[bin], (double)v[0]);
      atomicAdd(&s_c[bin], (unsigned long long)(v[1] + 0.5f));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    if (s_c[i] != 0ull) {
      atomicAdd(ksum + i, s_k[i]);
      atomicAdd(cnt + i, s_c[i]);
    }
  }
}

// out arrays are [n_fields][kmax]; acc is [n_fields][nspec][kmax+1].
__global__ void pk_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ ksum,
                                   const unsigned long long* __restrict__ cnt, int n_fields, int nspec,
                                   int kmax, double inv_batch, double* __restrict__ k_mean,
                                   double* __restrict__ p0, double* __restrict__ p1,
                                   double* __restrict__ p2, int64_t* __restrict__ n_modes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_fields * kmax) return;
  const int f = i / kmax, b = i % kmax + 1;
  const double n = (double)cnt[b];
  const int nb = kmax + 1;
  k_mean[i] = ksum[b] / n;
  n_modes[i] = (int64_t)cnt[b];
  double* outs[3] = {p0, p1, p2};
  for (int s = 0; s < nspec; ++s) outs[s][i] = acc[((int64_t)f * nspec + s) * nb + b] * inv_batch / n;
}

// ---- cuFFT plan cache ----------------------------------------------------------------------
static std::mutex g_plan_mutex;
static std::map<std::tuple<int, int, int, int, int>, cufftHandle> g_plans;

static int get_plan(int dev, int n0, int n1, int n2, int count, cufftHandle* out) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  auto key = std::make_tuple(dev, n0, n1, n2, count);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    *out = it->second;
    return VDM_OK;
  }
  cufftHandle plan;
  int dims3[3] = {n0, n1, n2};
  int dims2[2] = {n1, n2};
  cufftResult r = (n0 == 1)
                      ? cufftPlanMany(&plan, 2, dims2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, count)
                      : cufftPlanMany(&plan, 3, dims3, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, count);
  if (r != CUFFT_SUCCESS) {
    set_error("cufftPlanMany(%d,%d,%d x%d) failed: %d", n0, n1, n2, count, (int)r);
    return VDM_E_CUFFT;
  }
  g_plans[key] = plan;
  *out = plan;
  return VDM_OK;
}

struct PkLayout {
  PkGeom g;
  int64_t count;        // transforms per input array
  size_t spec_bytes;    // one spectrum array
  size_t acc_off, ksum_off, cnt_off, total;
};

static PkLayout make_layout(int n_fields, int batch, int chan, int n0, int n1, int n2, int nspec,
                            int narrays) {
  PkLayout L;
  L.g.n0 = n0; L.g.n1 = n1; L.g.n2 = n2; L.g.n2h = n2 / 2 + 1;
  int kmin = n1 < n2 ? n1 : n2;
  if (n0 > 1 && n0 < kmin) kmin = n0;
  L.g.kmax = kmin / 2;
  L.g.last_even = (n2 % 2 == 0);
  L.g.modes = (int64_t)n0 * n1 * L.g.n2h;
  L.count = (int64_t)n_fields * batch * chan;
  L.spec_bytes = (size_t)L.count * L.g.modes * sizeof(float2);
  size_t off = (size_t)narrays * L.spec_bytes;
  off = (off + 255) & ~(size_t)255;
  L.acc_off = off;
  off += (size_t)n_fields * nspec * (L.g.kmax + 1) * sizeof(double);
  L.ksum_off = off;
  off += (size_t)(L.g.kmax + 1) * sizeof(double);
  L.cnt_off = off;
  off += (size_t)(L.g.kmax + 1) * sizeof(unsigned long long);
  L.total = off;
  return L;
}

static int pk_run(const float* f1, const float* f2, int n_fields, int batch, int chan, int n0, int n1,
                  int n2, void* work, size_t work_bytes, int nspec, double* k_mean, double* p0, double* p1,
                  double* p2, int64_t* n_modes, cudaStream_t stream) {
  VDM_CHECK_ARG(f1 && work && k_mean && p0 && n_modes, "vdm_pk: NULL pointer argument");
  VDM_CHECK_ARG(n_fields >= 1 && batch >= 1 && chan >= 1 && n0 >= 1 && n1 >= 2 && n2 >= 2,
                "vdm_pk: bad shape (%d,%d,%d,%d,%d,%d)", n_fields, batch, chan, n0, n1, n2);
  const int narrays = f2 ? 2 : 1;
  PkLayout L = make_layout(n_fields, batch, chan, n0, n1, n2, nspec, narrays);
  VDM_CHECK_ARG(work_bytes >= L.total, "vdm_pk: work buffer too small (%zu < %zu)", work_bytes, L.total);
  VDM_CHECK_ARG(L.count <= 0x7fffffff, "vdm_pk: too many transforms");
  VDM_CHECK_ARG(L.g.kmax >= 1 && L.g.kmax <= 1024, "vdm_pk: kmax %d out of range", L.g.kmax);

  int dev = 0;
  VDM_CHECK_CUDA(cudaGetDevice(&dev));
  cufftHandle plan;
  int rc = get_plan(dev, n0, n1, n2, (int)L.count, &plan);
  if (rc != VDM_OK) return rc;
  {
    // plan handles are shared: serialise stream binding + exec so concurrent callers do not race.
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    cufftResult r = cufftSetStream(plan, stream);
    if (r != CUFFT_SUCCESS) { set_error("cufftSetStream failed: %d", (int)r); return VDM_E_CUFFT; }
    char* base = static_cast<char*>(work);
    r = cufftExecR2C(plan, const_cast<float*>(f1), reinterpret_cast<cufftComplex*>(base));
    if (r != CUFFT_SUCCESS) { set_error("cufftExecR2C failed: %d", (int)r); return VDM_E_CUFFT; }
    if (f2) {
      r = cufftExecR2C(plan, const_cast<float*>(f2), reinterpret_cast<cufftComplex*>(base + L.spec_bytes));
      if (r != CUFFT_SUCCESS) { set_error("cufftExecR2C(2) failed: %d", (int)r); return VDM_E_CUFFT; }
    }
  }
  char* base = static_cast<char*>(work);
  const float2* X = reinterpret_cast<const float2*>(base);
  const float2* X2 = f2 ? reinterpret_cast<const float2*>(base + L.spec_bytes) : X;
  double* acc = reinterpret_cast<double*>(base + L.acc_off);
  double* ksum = reinterpret_cast<double*>(base + L.ksum_off);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(base + L.cnt_off);
  VDM_CHECK_CUDA(cudaMemsetAsync(base + L.acc_off, 0, L.total - L.acc_off, stream));

  const int nb = L.g.kmax + 1;
  const int threads = 256;
  // enough blocks to fill the 148 SMs a few times over, bounded so per-block flushes stay cheap
  int64_t want = (L.g.modes + threads * 4 - 1) / (threads * 4);
  int bx = (int)(want < 1 ? 1 : want);
  const int cap = (kNumSMs * 8 + n_fields - 1) / n_fields;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, n_fields);
  const int ntr = batch * chan;
  const size_t bin_smem = (size_t)(threads / 32) * nspec * nb * sizeof(double);     // private bins per warp
  VDM_CHECK_ARG(bin_smem <= 200 * 1024, "vdm_pk: grid too large for the shared-memory bins (kmax = %d)", L.g.kmax);
  if (nspec == 1) {
    if (bin_smem > 48 * 1024)
      VDM_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem));
    pk_bin_kernel<1><<<grid, threads, bin_smem, stream>>>(X, X2, L.g, ntr, acc);
  } else {
    if (bin_smem > 48 * 1024)
      VDM_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem));
    pk_bin_kernel<3><<<grid, threads, bin_smem, stream>>>(X, X2, L.g, ntr, acc);
  }
  VDM_CHECK_LAUNCH();
  int gx = (int)((L.g.modes + threads * 8 - 1) / (threads * 8));
  gx = gx < 1 ? 1 : (gx > kNumSMs * 4 ? kNumSMs * 4 : gx);
  pk_geometry_kernel<<<gx, threads, nb * (sizeof(double) + sizeof(unsigned long long)), stream>>>(L.g, ksum, cnt);
  VDM_CHECK_LAUNCH();
  const int total = n_fields * L.g.kmax;
  pk_finalize_kernel<<<ceil_div(total, 128), 128, 0, stream>>>(acc, ksum, cnt, n_fields, nspec, L.g.kmax,
                                                                1.0 / batch, k_mean, p0, p1, p2, n_modes);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}

}  // namespace vdm

extern "C" size_t vdm_pk_work_bytes(int n_fields, int batch, int chan, int n0, int n1, int n2, int cross) {
  if (n_fields < 1 || batch < 1 || chan < 1 || n0 < 1 || n1 < 2 || n2 < 2) return 0;
  return vdm::make_layout(n_fields, batch, chan, n0, n1, n2, 3, cross ? 2 : 1).total;
}

extern "C" int vdm_pk(const float* fields, const float* fields2, int n_fields, int batch, int chan, int n0,
                      int n1, int n2, void* work, size_t work_bytes, double* k_mean, double* p_mean,
                      int64_t* n_modes, void* stream) {
  return vdm::pk_run(fields, fields2, n_fields, batch, chan, n0, n1, n2, work, work_bytes, 1, k_mean, p_mean,
                     nullptr, nullptr, n_modes, (cudaStream_t)stream);
}

extern "C" int vdm_pk_cross3(const float* fields1, const float* fields2, int n_fields, int batch, int chan,
                             int n0, int n1, int n2, void* work, size_t work_bytes, double* k_mean,
                             double* p11, double* p22, double* p12, int64_t* n_modes, void* stream) {
  if (!fields2 || !p22 || !p12) {
    vdm::set_error("vdm_pk_cross3: NULL pointer argument");
    return VDM_E_BADARG;
  }
  return vdm::pk_run(fields1, fields2, n_fields, batch, chan, n0, n1, n2, work, work_bytes, 3, k_mean, p11, p22,
                     p12, n_modes, (cudaStream_t)stream);
}
