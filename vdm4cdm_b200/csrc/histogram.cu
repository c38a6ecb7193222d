// log-PDF histograms of the generated mass fields: calc_SS.py:51-65 (get_logpdf_3d / get_logpdf_2d):
//     logfields = torch.log10(fields + 1);  np.histogram(logfields[i].flatten(), bins=np.linspace(lo, hi, nbins + 1))
// One pass over the field (4 B/voxel, HBM-bound), one histogram per field: per-block shared-memory bins with
// warp-aggregated increments (density fields are sharply peaked: most lanes of a warp hit the same few bins), then
// one 64-bit global atomic per bin and block.
//
// Bit-compatibility with the reference: the reference takes the log in fp32 (ATen / glibc log10f, correctly rounded
// in practice) and bins the fp32 value against fp64 edges lo + i*(hi-lo)/nbins with numpy's rule (left-closed bins,
// the last one closed on the right).  CUDA's fast log10f can differ in the last bit, which moves values sitting on a
// bin edge; so the log is taken in fp64 and rounded to fp32 (== the correctly rounded fp32 result), and the bin search
// runs in fp64 with numpy's edge correction.
#include "common.cuh"

namespace vdm {

constexpr int kHistMaxBins = 1024;

__global__ void __launch_bounds__(256)
log_histogram_kernel(const float* __restrict__ fields, long long voxels, float add, double lo, double hi, int nbins,
                     unsigned long long* __restrict__ counts) {
  __shared__ unsigned int s_bins[kHistMaxBins];
  for (int i = threadIdx.x; i < nbins; i += 256) s_bins[i] = 0u;
  __syncthreads();
  const float* f = fields + (long long)blockIdx.y * voxels;
  const double step = (hi - lo) / (double)nbins;
  const double norm = (double)nbins / (hi - lo);
  const long long n_iter = (voxels + (long long)gridDim.x * 256 - 1) / ((long long)gridDim.x * 256);
  for (long long it = 0; it < n_iter; ++it) {
    const long long i = (it * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
    int bin = -1;
    if (i < voxels) {
      const float xs = f[i] + add;                                   // fp32 sum, as torch computes fields + 1
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
      if (x >= lo && x <= hi) {
        int b = (int)((x - lo) * norm);
        if (b >= nbins) b = nbins - 1;                               // x == hi: last bin is closed on the right
        // numpy's correction against the actual edges
        if (x < lo + (double)b * step) --b;
        else if (b + 1 < nbins && x >= lo + (double)(b + 1) * step) ++b;
        bin = b;
      }
    }
    // warp-aggregated increment: lanes with equal bins elect one to add their count
    const unsigned mask = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && (int)(__ffs(mask) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_bins[bin], (unsigned)__popc(mask));
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
  long long blocks = (voxels + 256 * 8 - 1) / (256 * 8);
  const long long cap = (kNumSMs * 8 + n_fields - 1) / n_fields;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  log_histogram_kernel<<<dim3((unsigned)blocks, (unsigned)n_fields), 256, 0, (cudaStream_t)stream>>>(
      fields, voxels, add, lo, hi, nbins, reinterpret_cast<unsigned long long*>(counts));
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}
