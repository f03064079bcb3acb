"""CPU oracle for the raster-image kernels either side of the denoising path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product.

numpy restatement of three pieces of the reference, each pinned against the reference's own functions run in the build
container (tests/golden/make_raster_golden.py -> tests/golden/raster_*.npz, tests/test_oracle_raster.py):

* ``image_to_sample``   Image_Dataset.__getitem__  (DriveSceneGen/utils/datasets/dataset.py:20-23,44-47)
* ``gray_mask``         get_gray_image             (DriveSceneGen/vectorization/utils/image_utils.py:13-42)
* ``agent_threshold``   extract_agents, head       (DriveSceneGen/vectorization/direct/extract_vehicles.py:136-148)
* ``resize_to_sample``  Image_Dataset.__getitem__ with its live Resize((H, W), antialias=False) and the .pkl branch
                        (dataset.py:20-23,38-47); pinned by tests/golden/make_resize_golden.py -> resize_golden.npz

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


# ---------------------------------------------------------------------------------------------------------------------
# Resize((H, W), antialias=False) of Image_Dataset (DriveSceneGen/utils/datasets/dataset.py:20-23): torchvision's
# tensor resize = torch.nn.functional.interpolate(mode="bilinear", align_corners=False) = ATen's CPU
# upsample_bilinear2d (UpSampleKernel.cpp, generic N-d linear kernel), restated here bit for bit as the container's torch
# 2.11 / AVX-512 build evaluates it (pinned by tests/golden/resize_golden.npz, which torchvision itself produced):
#   scale  = float(in) / float(out)
#   src    = max(fma(scale, i + 0.5, -0.5), 0);  i0 = min(int(src), in - 1);  i1 = i0 + (i0 < in - 1)
#   l1     = clamp(src - i0, 0, 1);  l0 = 1 - l1
#   1-D    : lerp(a, b) = fma(a, l0, b * l1)
#   2-D    : lerp_y(lerp_x(row y0), lerp_x(row y1))
def fma32(a, b, c) -> np.ndarray:
    """Exactly rounded float32 fma(a, b, c) for float32 arrays: the product is exact in float64; the sum is rounded to
    odd in float64 (TwoSum error term), which makes the final rounding to float32 the correct single rounding."""
    a = np.asarray(a, np.float32).astype(np.float64)
    b = np.asarray(b, np.float32).astype(np.float64)
    c = np.asarray(c, np.float32).astype(np.float64)
    a, b, c = np.broadcast_arrays(a, b, c)
    p = a * b
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    s = np.ascontiguousarray(s)
    even = (s.view(np.int64) & 1) == 0
    nudge = (err != 0) & even
    s = np.where(nudge, np.nextafter(s, np.where(err > 0, np.inf, -np.inf)), s)
    return s.astype(np.float32)


def bilinear_taps(in_size: int, out_size: int):
    """(i0, i1, l0, l1) of ATen's compute_indices_weights for linear interpolation, align_corners=False."""
    f32 = np.float32
    if in_size == out_size:   # ATen: scale 1 -> plain copy (weights 1, 0)
        i0 = np.arange(out_size, dtype=np.int64)
        return i0, i0.copy(), np.ones(out_size, f32), np.zeros(out_size, f32)
    scale = f32(in_size) / f32(out_size)
    i = np.arange(out_size, dtype=f32)
    src = np.maximum(fma32(np.full_like(i, scale), i + f32(0.5), np.full_like(i, f32(-0.5))), f32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    l1 = np.minimum(np.maximum(src - i0.astype(f32), f32(0)), f32(1)).astype(f32)
    i1 = np.where(i0 < in_size - 1, i0 + 1, i0)
    return i0, i1, (f32(1) - l1).astype(f32), l1


def resize_bilinear(x: np.ndarray, out_h: int, out_w: int, mode: int = 0) -> np.ndarray:
    """float32 [..., h, w] -> float32 [..., out_h, out_w] (torchvision Resize(antialias=False) on a float tensor).

    mode 0: ATen's generic N-d kernel (multi-threaded host, out_h + out_w > 128 — the reference's 512^2 -> 256^2);
    mode 1: ATen's channels-last kernel (single-threaded host with 3 channels, or out_h + out_w <= 128):
            out = fma(d, w11, fma(c, w10, fma(a, w00, b * w01))), w_yx = l_y * l_x."""
    x = np.asarray(x, np.float32)
    y0, y1, wy0, wy1 = bilinear_taps(x.shape[-2], out_h)
    x0, x1, wx0, wx1 = bilinear_taps(x.shape[-1], out_w)
    a, b = x[..., y0, :][..., x0], x[..., y0, :][..., x1]
    c, d = x[..., y1, :][..., x0], x[..., y1, :][..., x1]
    if mode == 1:
        f32 = np.float32
        w00, w01 = (wy0[:, None] * wx0).astype(f32), (wy0[:, None] * wx1).astype(f32)
        w10, w11 = (wy1[:, None] * wx0).astype(f32), (wy1[:, None] * wx1).astype(f32)
        return fma32(d, w11, fma32(c, w10, fma32(a, w00, (b * w01).astype(f32))))

    def lerp(p, q, l0, l1):
        return fma32(p, l0, (q * l1).astype(np.float32))

    return lerp(lerp(a, b, wx0, wx1), lerp(c, d, wx0, wx1), wy0[:, None], wy1[:, None])


def resize_to_sample(img: np.ndarray, out_h: int, out_w: int, c_out: int | None = None, mode: int = 0) -> np.ndarray:
    """Image_Dataset.__getitem__ with a live Resize (dataset.py:20-23,38-47): [n,h,w,c] uint8 (ToTensor: /255) or float32
    (the .pkl branch) -> Resize((out_h, out_w), antialias=False) -> Normalize([0.5],[0.5]) -> float32 [n,c_out,H,W]."""
    img = np.asarray(img)
    assert img.ndim == 4 and img.dtype in (np.uint8, np.float32)
    c_out = img.shape[3] if c_out is None else c_out
    x = img[..., :c_out].astype(np.float32)
    if img.dtype == np.uint8:
        x = x / np.float32(255.0)
    x = np.ascontiguousarray(x.transpose(0, 3, 1, 2))
    if (out_h, out_w) != x.shape[-2:]:
        x = resize_bilinear(x, out_h, out_w, mode)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)
