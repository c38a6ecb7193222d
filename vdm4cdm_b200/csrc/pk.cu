// P(k) / r(k): cuFFT R2C followed by a single-pass k-shell binning kernel.
//
// Stands in for power() at src/utils.py:16-83 of the reference (pk :85-102, get_ccs :110-128):
// rfftn -> X conj(X2) -> batch mean / channel sum -> ceil(|k|) bins with Hermitian weights.
// The reference builds ~10 full-size temporaries and calls torch.bincount three times; here the
// wave number is recomputed from the mode's coordinates, the 32 consecutive modes of a warp trip are
// summed per run of equal bins (one ballot + a five-round segmented shuffle scan) and the run tails go to
// fp64 bins private to the warp with plain read-modify-writes; blocks flush once with fp64 global atomics.  k-mean and mode counts depend on the grid only and
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

// Position of a thread in the half spectrum, advanced by a fixed stride without divisions: m = (i0 * n1 + i1) * n2h + i2.
struct ModeCursor {
  int i0, i1, i2;
  int ds, d1, d0;     // the stride in (i2, i1, i0) digits
  __device__ __forceinline__ void init(const PkGeom& g, int64_t m, int64_t stride) {
    i2 = (int)(m % g.n2h);
    const int64_t row = m / g.n2h;
    i1 = (int)(row % g.n1);
    i0 = (int)(row / g.n1);
    ds = (int)(stride % g.n2h);
    const int64_t drow = stride / g.n2h;
    d1 = (int)(drow % g.n1);
    d0 = (int)(drow / g.n1);
  }
  __device__ __forceinline__ void advance(const PkGeom& g) {
    i2 += ds;
    const int c2 = i2 >= g.n2h ? 1 : 0;
    i2 -= c2 ? g.n2h : 0;
    i1 += d1 + c2;                     // <= 2 * n1 - 1
    const int c1 = i1 >= g.n1 ? 1 : 0;
    i1 -= c1 ? g.n1 : 0;
    i0 += d0 + c1;
  }
  // bin = ceil(|k|) (src/utils.py:53-56), Hermitian weight (src/utils.py:59-66); bin = -1 outside 1..kmax
  __device__ __forceinline__ int bin(const PkGeom& g, float& weight, float& kmag) const {
    const int f0 = i0 > g.n0 / 2 ? i0 - g.n0 : i0;
    const int f1 = i1 > g.n1 / 2 ? i1 - g.n1 : i1;
    const int k2 = f0 * f0 + f1 * f1 + i2 * i2;
    kmag = sqrtf((float)k2);
    // ceil(|k|) in integer arithmetic: the smallest b with b^2 >= k2 (what ceil(sqrt(.)) of the reference gives with a
    // correctly rounded sqrt; independent of how this sqrtf rounds at perfect squares)
    int b = (int)kmag;
    while (b * b < k2) ++b;
    while (b > 0 && (b - 1) * (b - 1) >= k2) --b;
    weight = (i2 == 0 || (g.last_even && i2 == g.n2h - 1)) ? 1.f : 2.f;
    return (b >= 1 && b <= g.kmax) ? b : -1;
  }
};

constexpr int kPkThreads = 256, kPkWarps = kPkThreads / 32;
constexpr int kPkUnroll = 4;          // modes (independent 8-byte loads) in flight per thread

// Sum the values of every maximal run of equal `bin` among the 32 consecutive modes of a warp trip, without atomics: the
// run starts come from ONE ballot, a lane's run start from a count-leading-zeros of that mask, and the segmented inclusive
// scan is five rounds of NV shuffles + predicated adds ("is the source lane still inside my run").  Returns true on the
// last lane of a run, which then holds the run's sums.  (Round 1-2: a generic segmented scan that also shuffled the head
// flags, ~120 instructions per trip; R3: fp32 shared-memory atomicAdd instead -- which is NOT a native instruction on
// sm_100: ptxas emits an ATOMS.CAST.SPIN compare-and-swap loop, 0.11 ms per 16 fields of 128^3, R4e.)
template <int NV>
__device__ __forceinline__ bool warp_run_sums(int bin, float (&v)[NV]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(full, bin, 1);
  const unsigned starts = __ballot_sync(full, lane == 0 || prev != bin);
  const int my_start = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));      // highest run start at or below this lane
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float up = __shfl_up_sync(full, v[i], o);
      if (lane - o >= my_start) v[i] += up;
    }
  }
  return lane == 31 || ((starts >> (lane + 1)) & 1u);
}

// Add the run sums of a trip to the warp's PRIVATE fp64 bins.  Bins grow along a spectrum row, so a bin normally has one run
// (one tail lane) per trip and the tails do plain read-modify-writes.  A trip that spans a row boundary holds a descent; then
// the same bin can occur on both sides only if the largest bin after the boundary reaches the smallest before it (two integer
// warp reductions, REDUX) -- those trips, and trips with several boundaries (short rows of small grids), take the
// compare-and-swap path.  (match.any over the tail lanes instead of this test cost 0.08 ms per 16 fields: R4f.)
template <int NV>
__device__ __forceinline__ void warp_bins_add(double* my, int nb, int bin, bool tail, const float (&v)[NV]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int key = bin >= 1 ? bin : 0x7fffffff;            // modes outside 1..kmax: neutral for the order test
  const int prev = __shfl_up_sync(full, key, 1);
  const unsigned desc = __ballot_sync(full, lane > 0 && prev > key);
  bool dup = false;
  if (desc != 0u) {                                        // warp-uniform
    if ((desc & (desc - 1u)) != 0u) {
      dup = true;
    } else {
      const int p = __ffs(desc) - 1;                       // first lane after the boundary
      const int max_after = __reduce_max_sync(full, (lane >= p && bin >= 1) ? bin : -1);
      const int min_before = __reduce_min_sync(full, (lane < p && bin >= 1) ? bin : 0x7fffffff);
      dup = max_after >= min_before;
    }
  }
  if (tail && bin >= 1) {
    if (!dup) {
#pragma unroll
      for (int i = 0; i < NV; ++i) my[i * nb + bin] += (double)v[i];
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) atomicAdd(&my[i * nb + bin], (double)v[i]);
    }
  }
  __syncwarp();
}

// acc: double [n_fields][NSPEC][kmax+1], zero-initialised.
// NSPEC == 1: Re(X conj(X2)) (X2 may alias X).  NSPEC == 3: |X|^2, |X2|^2, Re(X conj(X2)).
//
// One mode per thread and trip, kPkUnroll trips' loads in flight, coordinates advanced without divisions (ModeCursor); the
// 32 consecutive modes of a warp trip are reduced per run of equal bins (warp_run_sums) and the run tails go to fp64 bins
// private to the warp; blocks fold their warps' bins in a fixed order and flush once with fp64 global atomics.
// AUTO: X2 aliases X (auto-spectrum: one load per mode).  SINGLE: one transform per field (no batch / channel loop).  Mode
// indices are 32-bit inside a field (checked by the host).  (R4g: with a run-time transform loop, two loads per mode even
// for the auto-spectrum and 64-bit index arithmetic the loop body was ~500 instructions per mode and the pass took 0.12 ms
// per 16 fields of 128^3 whatever the reduction scheme -- it was issue-bound on its own addressing.)
template <int NSPEC, bool AUTO, bool SINGLE>
__global__ void __launch_bounds__(kPkThreads)
pk_bin_kernel(const float2* __restrict__ X, const float2* __restrict__ X2, PkGeom g, int n_transforms,
              double* __restrict__ acc) {
  extern __shared__ double s_acc[];   // [warp][NSPEC][kmax+1]
  const int nb = g.kmax + 1, n = NSPEC * nb;
  for (int i = threadIdx.x; i < kPkWarps * n; i += kPkThreads) s_acc[i] = 0.0;
  __syncthreads();
  double* my = s_acc + (threadIdx.x >> 5) * n;

  const int field = blockIdx.y;
  const float2* x = X + (int64_t)field * n_transforms * g.modes;
  const float2* x2 = X2 + (int64_t)field * n_transforms * g.modes;
  const int modes = (int)g.modes;
  const int stride = (int)gridDim.x * kPkThreads;
  int m = (int)blockIdx.x * kPkThreads + (int)threadIdx.x;
  ModeCursor cur;
  cur.init(g, m, stride);
  const int trips = (modes + stride * kPkUnroll - 1) / (stride * kPkUnroll);   // the same for every thread
  for (int it = 0; it < trips; ++it) {
    int bins[kPkUnroll];
    float v[kPkUnroll][NSPEC];
#pragma unroll
    for (int u = 0; u < kPkUnroll; ++u) {
      float w, kmag;
      bins[u] = (m < modes) ? cur.bin(g, w, kmag) : -1;
#pragma unroll
      for (int i = 0; i < NSPEC; ++i) v[u][i] = 0.f;
      if (bins[u] >= 1) {
        const int nt = SINGLE ? 1 : n_transforms;
        const float2* xa = x + m;
        const float2* xc = x2 + m;
        for (int t = 0; t < nt; ++t, xa += modes, xc += modes) {
          const float2 a = __ldg(xa);
          if constexpr (NSPEC == 1) {
            if constexpr (AUTO) {
              v[u][0] += a.x * a.x + a.y * a.y;
            } else {
              const float2 c = __ldg(xc);
              v[u][0] += a.x * c.x + a.y * c.y;
            }
          } else {
            const float2 c = __ldg(xc);
            v[u][0] += a.x * a.x + a.y * a.y;
            v[u][NSPEC - 2] += c.x * c.x + c.y * c.y;
            v[u][NSPEC - 1] += a.x * c.x + a.y * c.y;
          }
        }
#pragma unroll
        for (int i = 0; i < NSPEC; ++i) v[u][i] *= w;
      }
      m += stride;
      cur.advance(g);
    }
#pragma unroll
    for (int u = 0; u < kPkUnroll; ++u) {
      const bool tail = warp_run_sums<NSPEC>(bins[u], v[u]);
      warp_bins_add<NSPEC>(my, nb, bins[u], tail, v[u]);
    }
  }
  __syncthreads();
  double* out = acc + (int64_t)field * n;
  for (int i = threadIdx.x; i < n; i += kPkThreads) {
    double t = 0.0;
#pragma unroll
    for (int wv = 0; wv < kPkWarps; ++wv) t += s_acc[wv * n + i];
    if (t != 0.0) atomicAdd(out + i, t);
  }
}

// geometry: ksum[bin] += w*|k| (double), cnt[bin] += w (unsigned long long). No memory reads; same traversal.
__global__ void __launch_bounds__(kPkThreads)
pk_geometry_kernel(PkGeom g, double* __restrict__ ksum, unsigned long long* __restrict__ cnt) {
  extern __shared__ double s_acc[];   // [warp][kmax+1] fp64 k sums, then [kmax+1] counts
  const int nb = g.kmax + 1;
  unsigned int* s_c = reinterpret_cast<unsigned int*>(s_acc + kPkWarps * nb);
  for (int i = threadIdx.x; i < kPkWarps * nb; i += kPkThreads) s_acc[i] = 0.0;
  for (int i = threadIdx.x; i < nb; i += kPkThreads) s_c[i] = 0u;
  __syncthreads();
  double* my = s_acc + (threadIdx.x >> 5) * nb;
  const int64_t stride = (int64_t)gridDim.x * kPkThreads;
  int64_t m = (int64_t)blockIdx.x * kPkThreads + threadIdx.x;
  ModeCursor cur;
  cur.init(g, m, stride);
  const int trips = (int)((g.modes + stride - 1) / stride);
  for (int it = 0; it < trips; ++it) {
    int b = -1;
    float v[2] = {0.f, 0.f};
    if (m < g.modes) {
      float w, kmag;
      b = cur.bin(g, w, kmag);
      if (b >= 1) {
        v[0] = kmag * w;
        v[1] = w;                                  // <= 64 per run: exact in fp32
      }
    }
    const bool tail = warp_run_sums<2>(b, v);
    const float one[1] = {v[0]};
    warp_bins_add<1>(my, nb, b, tail, one);
    if (tail && b >= 1) atomicAdd(&s_c[b], (unsigned int)(v[1] + 0.5f));      // integer shared atomics are native
    m += stride;
    cur.advance(g);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += kPkThreads) {
    if (s_c[i] != 0u) {
      double t = 0.0;
#pragma unroll
      for (int wv = 0; wv < kPkWarps; ++wv) t += s_acc[wv * nb + i];
      atomicAdd(ksum + i, t);
      atomicAdd(cnt + i, (unsigned long long)s_c[i]);
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
  const int threads = kPkThreads;
  // enough blocks to fill the 148 SMs (8 resident blocks each), bounded so per-block flushes stay cheap
  int64_t want = (L.g.modes + threads * kPkUnroll - 1) / (threads * kPkUnroll);
  int bx = (int)(want < 1 ? 1 : want);
  const int cap = (kNumSMs * 4 + n_fields - 1) / n_fields;   // 64 registers x 256 threads: four blocks per SM
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, n_fields);
  const int ntr = batch * chan;
  const size_t bin_smem = (size_t)kPkWarps * nspec * nb * sizeof(double);     // fp64 bins private to each warp
  VDM_CHECK_ARG(bin_smem <= 200 * 1024, "vdm_pk: grid too large for the shared-memory bins (kmax = %d)", L.g.kmax);
  VDM_CHECK_ARG(L.g.modes + (int64_t)bx * threads * kPkUnroll < ((int64_t)1 << 31), "vdm_pk: field too large for 32-bit mode indices");
  const bool aut = f2 == nullptr, single = ntr == 1;
#define VDM_PK_LAUNCH(NS, AU, SI)                                                                                        \
  {                                                                                                                      \
    if (bin_smem > 48 * 1024)                                                                                            \
      VDM_CHECK_CUDA(cudaFuncSetAttribute(pk_bin_kernel<NS, AU, SI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem)); \
    pk_bin_kernel<NS, AU, SI><<<grid, threads, bin_smem, stream>>>(X, X2, L.g, ntr, acc);                                 \
  }
  if (nspec == 1) {
    if (aut && single) VDM_PK_LAUNCH(1, true, true)
    else if (aut) VDM_PK_LAUNCH(1, true, false)
    else if (single) VDM_PK_LAUNCH(1, false, true)
    else VDM_PK_LAUNCH(1, false, false)
  } else {
    if (single) VDM_PK_LAUNCH(3, false, true)
    else VDM_PK_LAUNCH(3, false, false)
  }
#undef VDM_PK_LAUNCH
  VDM_CHECK_LAUNCH();
  int gx = (int)((L.g.modes + threads * 8 - 1) / (threads * 8));
  gx = gx < 1 ? 1 : (gx > kNumSMs * 4 ? kNumSMs * 4 : gx);
  const size_t geo_smem = (size_t)nb * (kPkWarps * sizeof(double) + sizeof(unsigned int));
  pk_geometry_kernel<<<gx, threads, geo_smem, stream>>>(L.g, ksum, cnt);
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
