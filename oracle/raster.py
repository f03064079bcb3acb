"""CPU oracle for the raster-image kernels either side of the denoising path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product.

numpy restatement of three pieces of the reference, each pinned against the reference's own functions run in the build
container (tests/golden/make_raster_golden.py -> tests/golden/raster_*.npz, tests/test_oracle_raster.py):

* ``image_to_sample``   Image_Dataset.__getitem__  (DriveSceneGen/utils/datasets/dataset.py:20-23,44-47)
* ``gray_mask``         get_gray_image             (DriveSceneGen/vectorization/utils/image_utils.py:13-42)
* ``agent_threshold``   extract_agents, head       (DriveSceneGen/vectorization/direct/extract_vehicles.py:136-148)

``np.histogram`` is the reference's own third-party call (numpy, requirements.txt) and is used as such.
"""
from __future__ import annotations

import numpy as np


def image_to_sample(img_u8: np.ndarray, c_out: int | None = None) -> np.ndarray:
    """uint8 [n,h,w,c] -> float32 [n,c_out,h,w]: ToTensor (x/255 in fp32) then Normalize([0.5],[0.5]) ((x-0.5)/0.5).

    dataset.py:44-47 (``transforms.ToTensor()(Image.open(f))``, ``self.normalize``); the Resize of dataset.py:21 is the
    identity for rasters stored at the model's size."""
    img = np.asarray(img_u8)
    assert img.dtype == np.uint8 and img.ndim == 4
    c_out = img.shape[3] if c_out is None else c_out
    x = img[..., :c_out].astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def gray_mask(img_u8: np.ndarray, thresh: float = 0.1):
    """uint8 [h,w,c>=3] -> (hist int64 [3,256], peaks int [3], mask uint8 [h,w]).  image_utils.py:13-42."""
    t = np.array(img_u8, dtype=float)                                   # image_utils.py:14
    chans = [t[:, :, k].flatten() / 255.0 for k in range(3)]            # :16-23 (divide then flatten: same values)
    hists, peaks = [], []
    for v in chans:
        h, bins = np.histogram(v, bins=256, range=(0, 1))               # :26-28
        hists.append(h)
        peaks.append(int(np.argmax(h)))                                 # :31-33 (bins[k] == k / 256 exactly)
    mx, my = peaks[0] / 256.0, peaks[1] / 256.0                          # :36-37
    near = (np.fabs(chans[0] - mx) <= thresh) & (np.fabs(chans[1] - my) <= thresh)   # combine_dx_dy, :6-10
    mask = np.where(near, 0, 255).astype(np.uint8).reshape(t.shape[:2])  # :40-41
    return np.stack(hists), np.array(peaks), mask


def agent_threshold(plane_f32: np.ndarray, thresh: int = 100) -> np.ndarray:
    """float32 [h,w] speed channel in [0,1] -> uint8 [h,w].  extract_vehicles.py:136-148.

    ``(image * 255).astype(np.uint8)`` truncates; ``cv2.cvtColor(BGR2GRAY)`` of three identical 8-bit channels v is
    ``(v*1868 + v*9617 + v*4899 + 8192) >> 14 == v``; ``cv2.threshold(gray, thresh, 255, 0)`` is ``> thresh``."""
    img = (np.asarray(plane_f32, dtype=np.float32) * 255).astype(np.uint8)
    gray = ((img.astype(np.int64) * (1868 + 9617 + 4899) + 8192) >> 14).astype(np.uint8)
    return np.where(gray > thresh, 255, 0).astype(np.uint8)
