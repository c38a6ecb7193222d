"""Oracle for the training-batch preparation (TEST INFRASTRUCTURE): numpy restatement of the reference's
per-sample pipeline, pinned to the reference's own classes by tests/golden/augment_golden.npz
(generator: oracle/make_golden_augment.py, which imports /root/reference/src/dataset/augmentation.py).

  Crop.__call__        src/dataset/augmentation.py:112-127   periodic window  i = arange(a - p0, a + c + p1) % s
  LogTransform         :8-21                                  log10(img + alpha)
  Normalize            :23-40                                 (img - mean) / std
  Flip.__call__        :49-59                                 torch.flip(img, 1 + axes)
  Permutate.__call__   :68-77                                 img.permute([0] + (1 + axes))
  AstroDataset.__getitem__  src/dataset/CAMELS_3D_dataset.py:53-74 (crop, float32, transform, return_func)
"""
import numpy as np


def crop_periodic(img, anchor, crop, pad=((0, 0), (0, 0), (0, 0))):
    """img: (C, S0, S1, S2).  Periodic window starting at anchor - pad[:,0] of extent crop + pad."""
    ndim = 3
    ind = [slice(None)]
    for d in range(ndim):
        a, c, (p0, p1), s = int(anchor[d]), int(crop[d]), pad[d], img.shape[1 + d]
        i = np.arange(a - p0, a + c + p1)
        i %= s
        ind.append(i.reshape((-1,) + (1,) * (ndim - d - 1)))
    return img[tuple(ind)]


def log_normalize(img, alpha, mean, std):
    return ((np.log10(img.astype(np.float32) + np.float32(alpha)) - np.float32(mean)) / np.float32(std)).astype(np.float32)


def flip(img, flip_mask):
    axes = [1 + d for d in range(3) if flip_mask[d]]
    return np.flip(img, axes) if axes else img


def permutate(img, perm):
    return np.transpose(img, [0] + [1 + int(a) for a in perm])


def prepare(img, anchor, crop, flip_mask, perm, alpha, mean, std):
    """The whole per-field pipeline in the reference's order: crop -> log -> normalise -> flip -> permute."""
    x = crop_periodic(img, anchor, crop)
    x = log_normalize(x, alpha, mean, std)
    x = flip(x, flip_mask)
    return np.ascontiguousarray(permutate(x, perm))
