// log-PDF histograms of the generated mass fields: calc_SS.py:51-65 (get_logpdf_3d / get_logpdf_2d):
//     logfields = torch.log10(fields + 1);  np.histogram(logfields[i].flatten(), bins=np.linspace(lo, hi, nbins + 1))
// One pass over the field (4 B/voxel, HBM-bound), one histogram per field: per-block shared-memory bins (native integer
// atomics; the lanes that share lane 0's bin are added once), then one 64-bit global atomic per bin and block.
//
// Bit-compatibility with the reference: the reference takes the log in fp32 (ATen / glibc log10f, correctly rounded
// in practice) and bins the fp32 value against fp64 edges lo + i*(hi-lo)/nbins with numpy's rule (left-closed bins,
// the last one closed on the right).  CUDA's fast log10f can differ in the last bit, which moves values sitting on a
// bin edge; so the log is taken in fp64 and rounded to fp32 (== the correctly rounded fp32 result), and the bin search
// runs in fp64 with numpy's edge correction.
#include "common.cuh"

namespace vdm {

constexpr int kHistMaxBins = 1024;

// bin of one value (numpy's rule), or -1 outside [lo, hi]
__device__ __forceinline__ int log_bin(float raw, float add, double lo, double hi, double step, double norm, int nbins) {
  const float xs = raw + add;                                      // fp32 sum, as torch computes fields + 1
  // The bin is decided by the correctly rounded fp32 log10 (fp64 log10, rounded).  log10f is within 2 ulp of it,
  // so unless it lands within 8 ulp of a bin edge or of the range ends it already decides the same bin
  // (r02l: fp64 log10 for every voxel made this kernel instruction-bound at 0.66 TB/s).
  const float xa = log10f(xs);
  double x = (double)xa;
  {
    const double pos = (x - lo) * norm;
    const double frac = pos - floor(pos);
    const double tol = 8.0 * 1.1920929e-07 * fabs(x) * norm + 1e-9;   // 8 ulp of x in bin units
    if (!(frac > tol && frac < 1.0 - tol) || !(x == x) || fabs(x) > 1e30) x = (double)(float)log10((double)xs);
  }
  if (!(x >= lo && x <= hi)) return -1;
  int b = (int)((x - lo) * norm);
  if (b >= nbins) b = nbins - 1;                                   // x == hi: last bin is closed on the right
  // numpy's correction against the actual edges
  if (x < lo + (double)b * step) --b;
  else if (b + 1 < nbins && x >= lo + (double)(b + 1) * step) ++b;
  return b;
}

// Increment shared-memory bins for one value per lane.  Integer shared atomics are native (ATOMS.ADD); density fields are
// sharply peaked, so the lanes that share lane 0's bin are counted with one ballot and added once, the others add 1 each.
// (Round 2: __match_any_sync over all lanes -- it iterates over the distinct keys of the warp and made this kernel
// instruction-bound at 0.85 TB/s, 0.158 ms per 16 fields of 128^3, R4e.)
__device__ __forceinline__ void warp_hist_add(unsigned int* s_bins, int bin) {
  const unsigned full = 0xffffffffu;
  const int b0 = __shfl_sync(full, bin, 0);
  const unsigned same = __ballot_sync(full, bin == b0);
  if ((threadIdx.x & 31) == 0) {
    if (b0 >= 0) atomicAdd(&s_bins[b0], (unsigned)__popc(same));
  } else if (bin >= 0 && bin != b0) {
    atomicAdd(&s_bins[bin], 1u);
  }
}

__global__ void __launch_bounds__(256)
log_histogram_kernel(const float* __restrict__ fields, long long voxels, float add, double lo, double hi, int nbins,
                     unsigned long long* __restrict__ counts) {
  __shared__ unsigned int s_bins[kHistMaxBins];
  for (int i = threadIdx.x; i < nbins; i += 256) s_bins[i] = 0u;
  __syncthreads();
  const float* f = fields + (long long)blockIdx.y * voxels;
  const double step = (hi - lo) / (double)nbins;
  const double norm = (double)nbins / (hi - lo);
  // four voxels per thread and trip as one 16-byte load when the field allows it (voxels % 4 == 0 keeps every field aligned)
  const bool vec = (voxels & 3) == 0 && (reinterpret_cast<uintptr_t>(fields) & 15) == 0;
  const long long n4 = vec ? voxels >> 2 : 0;
  const long long span = (long long)gridDim.x * 256;
  const long long n_iter4 = (n4 + span - 1) / span;
  const float4* f4 = reinterpret_cast<const float4*>(f);
  for (long long it = 0; it < n_iter4; ++it) {
    const long long i = (it * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
    int b[4] = {-1, -1, -1, -1};
    if (i < n4) {
      const float4 v = __ldg(f4 + i);
      b[0] = log_bin(v.x, add, lo, hi, step, norm, nbins);
      b[1] = log_bin(v.y, add, lo, hi, step, norm, nbins);
      b[2] = log_bin(v.z, add, lo, hi, step, norm, nbins);
      b[3] = log_bin(v.w, add, lo, hi, step, norm, nbins);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) warp_hist_add(s_bins, b[k]);
  }
  const long long first = n4 << 2;                               // scalar remainder (or everything, unaligned fields)
  const long long n_iter = (voxels - first + span - 1) / span;
  for (long long it = 0; it < n_iter; ++it) {
    const long long i = first + (it * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
    const int bin = i < voxels ? log_bin(f[i], add, lo, hi, step, norm, nbins) : -1;
    warp_hist_add(s_bins, bin);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += 256)
    if (s_bins[i]) atomicAdd(counts + (long long)blockIdx.y * nbins + i, (unsigned long long)s_bins[i]);
}

}  // namespace vdm

using namespace vdm;

extern "C" int vdm_log_histogram(const float* fields, int n_fields, int64_t voxels, float add, double lo, double hi,
                                 int nbins, int64_t* counts, void* stream) {
  VDM_CHECK_ARG(fields && counts && n_fields >= 1 && voxels >= 1, "vdm_log_histogram: bad argument");
  VDM_CHECK_ARG(nbins >= 1 && nbins <= kHistMaxBins && hi > lo, "vdm_log_histogram: need 1 <= nbins <= %d and hi > lo", kHistMaxBins);
  VDM_CHECK_ARG(n_fields <= 65535, "vdm_log_histogram: too many fields");
  long long blocks = (voxels + 256 * 16 - 1) / (256 * 16);
  const long long cap = (kNumSMs * 8 + n_fields - 1) / n_fields;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  log_histogram_kernel<<<dim3((unsigned)blocks, (unsigned)n_fields), 256, 0, (cudaStream_t)stream>>>(
      fields, voxels, add, lo, hi, nbins, reinterpret_cast<unsigned long long*>(counts));
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}
