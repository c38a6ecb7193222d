// Training-batch preparation on the GPU: periodic crop + log-normalisation + flip + axis permutation in ONE gather.
//
// Stands in for the per-sample numpy/torch pipeline of the reference's DataLoader workers
// (AstroDataset.__getitem__: src/dataset/CAMELS_3D_dataset.py:53-74 -> Crop: src/dataset/augmentation.py:80-127,
// LogTransform/Normalize :8-40, Flip :43-59, Permutate :62-77; composed at CAMELS_3D_dataset.py:97-103).  With the
// training step at ~19 ms on one B200 (8 GPUs: ~850 samples/s) the 16-worker CPU pipeline becomes the bottleneck
// (SURVEY.md section 8f row 1); here the raw simulation boxes stay resident in HBM and a batch element costs one
// pass: out[o] = (log10(raw[src(o)] + alpha) - mean) / std with
//     p = inverse-permuted o,  f = flipped p,  src_a = (anchor_a + f_a) mod S      (all three axes).
// HBM-bound: 4 B read + 4 B written per output voxel (reads are strided when the permutation moves the fast axis;
// a 128^3 crop is 8 MB, three orders of magnitude below the training step's traffic).
#include "common.cuh"

namespace vdm {

struct AugParams {
  int S[3];        // full box (d, h, w)
  int n[3];        // OUTPUT grid (after permutation)
  int anchor[3];   // crop start per SOURCE axis (may be negative / >= S: periodic)
  int flip[3];     // flip per CROPPED axis (before permutation)
  int perm[3];     // output axis d takes cropped axis perm[d]   (img.permute([0] + (1 + axes)))
  int crop[3];     // crop extent per cropped axis (= n[inv perm])
  float alpha, mean, std;
  int do_log;
};

__global__ void __launch_bounds__(256)
augment_crop_kernel(const float* __restrict__ in, float* __restrict__ out, const AugParams p, long long total) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long v = i;
    int o[3];
    o[2] = (int)(v % p.n[2]); v /= p.n[2];
    o[1] = (int)(v % p.n[1]);
    o[0] = (int)(v / p.n[1]);
    int src[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int a = p.perm[d];                       // cropped axis that lands on output axis d
      int f = o[d];
      if (p.flip[a]) f = p.crop[a] - 1 - f;
      int s = (p.anchor[a] + f) % p.S[a];
      if (s < 0) s += p.S[a];
      src[a] = s;
    }
    float x = in[((long long)src[0] * p.S[1] + src[1]) * p.S[2] + src[2]];
    // fp64 log rounded to fp32 == the correctly rounded fp32 log10 the reference's CPU path produces; then the same
    // fp32 subtract and divide as torchvision's normalize
    if (p.do_log) x = ((float)log10((double)(x + p.alpha)) - p.mean) / p.std;
    out[i] = x;
  }
}

}  // namespace vdm

using namespace vdm;

extern "C" int vdm_augment_crop(const float* in, float* out, const int32_t* full_size, const int32_t* crop_size,
                                const int32_t* anchor, const int32_t* flip, const int32_t* perm, float alpha, float mean,
                                float std, int do_log, void* stream) {
  VDM_CHECK_ARG(in && out && full_size && crop_size && anchor && flip && perm, "vdm_augment_crop: NULL argument");
  AugParams p;
  bool seen[3] = {false, false, false};
  for (int d = 0; d < 3; ++d) {
    VDM_CHECK_ARG(full_size[d] >= 1 && crop_size[d] >= 1, "vdm_augment_crop: bad size on axis %d", d);
    VDM_CHECK_ARG(perm[d] >= 0 && perm[d] < 3 && !seen[perm[d]], "vdm_augment_crop: perm is not a permutation of (0,1,2)");
    seen[perm[d]] = true;
    p.S[d] = full_size[d]; p.crop[d] = crop_size[d]; p.anchor[d] = anchor[d]; p.flip[d] = flip[d] ? 1 : 0; p.perm[d] = perm[d];
  }
  for (int d = 0; d < 3; ++d) p.n[d] = p.crop[p.perm[d]];
  VDM_CHECK_ARG(!do_log || std != 0.f, "vdm_augment_crop: std must be non-zero");
  p.alpha = alpha; p.mean = mean; p.std = do_log ? std : 1.0f; p.do_log = do_log;
  const long long total = (long long)p.n[0] * p.n[1] * p.n[2];
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  augment_crop_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, p, total);
  VDM_CHECK_LAUNCH();
  return VDM_OK;
}
