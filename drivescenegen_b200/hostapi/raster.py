"""Host side of the raster-image kernels (libdsg_b200: dsg_image_to_sample, dsg_gray_mask, dsg_agent_threshold).

Mirrors three per-pixel pieces of the reference that sit either side of the denoising path, for tensors that are (or are
about to be) on the device:

* :func:`image_to_sample` — ``Image_Dataset.__getitem__`` arithmetic (DriveSceneGen/utils/datasets/dataset.py:20-23,44-47)
  after a uint8 host->device copy (4x fewer PCIe bytes than the fp32 batch the reference DataLoader ships);
* :func:`get_gray_image` / :func:`gray_masks` — ``image_utils.get_gray_image``
  (DriveSceneGen/vectorization/utils/image_utils.py:13-42); the result can be handed to the reference's
  ``extract_polylines_from_img(img_color, img_gray=...)`` (image_to_polylines.py:611, image_to_vectors_graph.py:410);
* :func:`agent_threshold` — the thresholded speed channel ``extract_agents`` feeds to ``cv2.findContours``
  (DriveSceneGen/vectorization/direct/extract_vehicles.py:136-148).

No CPU path: CPU tensors are uploaded, the kernels always run on the device.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import DsgError, check


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise DsgError("dsg_b200 raster kernels need a CUDA device (there is no CPU path)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _as_u8_batch(images, device=None) -> torch.Tensor:
    """PIL image / numpy array / tensor, [h,w,c] or [n,h,w,c] uint8 -> contiguous CUDA uint8 [n,h,w,c]."""
    if not isinstance(images, torch.Tensor):
        images = torch.from_numpy(np.ascontiguousarray(np.asarray(images)))
    if images.dtype != torch.uint8:
        raise ValueError(f"expected uint8 images, got {images.dtype}")
    if images.dim() == 3:
        images = images[None]
    if images.dim() != 4:
        raise ValueError(f"expected [h,w,c] or [n,h,w,c] images, got shape {tuple(images.shape)}")
    if not images.is_cuda:
        images = images.to(_device(device), non_blocking=True)
    return images.contiguous()


def image_to_sample(images, channels: int = 3, out: torch.Tensor = None, device=None) -> torch.Tensor:
    """uint8 rasters [n,h,w,c] (or one [h,w,c]) -> normalised fp32 samples [n,channels,h,w] in [-1, 1]."""
    img = _as_u8_batch(images, device)
    n, h, w, c = img.shape
    if not 1 <= channels <= c:
        raise ValueError(f"channels={channels} but the images have {c}")
    if out is None:
        out = torch.empty((n, channels, h, w), dtype=torch.float32, device=img.device)
    elif out.shape != (n, channels, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != img.device:
        raise ValueError("out must be a contiguous fp32 tensor [n,channels,h,w] on the images' device")
    check(_lib.load().dsg_image_to_sample(img.data_ptr(), out.data_ptr(), n, h, w, c, channels, _stream(img.device)),
          "dsg_image_to_sample")
    return out


def gray_masks(images, thresh: float = 0.1, want_gray3: bool = False, device=None):
    """Batched ``get_gray_image``: uint8 [n,h,w,c] -> (mask uint8 [n,h,w] or [n,h,w,3], peaks int32 [n,3], hist int32
    [n,3,256] holding uint32 counts), all on the device."""
    img = _as_u8_batch(images, device)
    n, h, w, c = img.shape
    if c not in (3, 4):
        raise ValueError(f"expected RGB or RGBA rasters, got {c} channels")
    dev = img.device
    hist = torch.empty((n, 3, 256), dtype=torch.int32, device=dev)
    peaks = torch.empty((n, 3), dtype=torch.int32, device=dev)
    mask = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    gray3 = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev) if want_gray3 else None
    check(_lib.load().dsg_gray_mask(img.data_ptr(), hist.data_ptr(), peaks.data_ptr(), mask.data_ptr(),
                                    gray3.data_ptr() if want_gray3 else None, n, h, w, c, float(thresh), _stream(dev)),
          "dsg_gray_mask")
    return (gray3 if want_gray3 else mask), peaks, hist


def get_gray_image(img_color, plot: bool = False):
    """Drop-in for ``image_utils.get_gray_image(img_color, plot=False)``: PIL image in, 3-channel PIL mask out."""
    if plot:
        raise ValueError("plot=True is a matplotlib debugging aid of the reference and is not provided")
    from PIL import Image
    arr = np.asarray(img_color)
    if arr.ndim != 3 or arr.shape[2] < 3:
        raise ValueError(f"expected a colour image, got array of shape {arr.shape}")
    gray3, _, _ = gray_masks(arr, want_gray3=True)
    return Image.fromarray(gray3[0].cpu().numpy())


def agent_threshold(raw_img: torch.Tensor, thresh: int = 100, channel: int = 2) -> torch.Tensor:
    """fp32 CHW image(s) in [0,1] (``ToTensor`` output, [3,h,w] or [n,3,h,w]) -> uint8 [n,h,w] blob mask of the speed
    channel, ready for ``cv2.findContours`` on the host."""
    if raw_img.dtype != torch.float32:
        raise ValueError(f"expected a float32 image tensor, got {raw_img.dtype}")
    x = raw_img[None] if raw_img.dim() == 3 else raw_img
    if x.dim() != 4 or not 0 <= channel < x.shape[1]:
        raise ValueError(f"expected [c,h,w] or [n,c,h,w] with c > {channel}, got {tuple(raw_img.shape)}")
    if not x.is_cuda:
        x = x.to(_device(), non_blocking=True)
    x = x.contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, h, w), dtype=torch.uint8, device=x.device)
    plane = x.data_ptr() + channel * h * w * 4
    check(_lib.load().dsg_agent_threshold(plane, c * h * w, out.data_ptr(), n, h * w, int(thresh), _stream(x.device)),
          "dsg_agent_threshold")
    return out


class RasterDataset(torch.utils.data.Dataset):
    """``Image_Dataset`` (DriveSceneGen/utils/datasets/dataset.py:15-50) with the arithmetic left for the device.

    Same constructor argument (``config.dataset_name`` glob, ``patterns_size_height/width``), same ``data_list`` /
    ``remove_sample``; ``__getitem__`` only decodes the PNG and returns the uint8 ``[h, w, c]`` raster.  A DataLoader over
    it collates uint8 ``[b, h, w, c]`` batches; ``Accelerator.prepare`` (ShardedDataLoader) ships those bytes to the GPU
    and runs ``dsg_image_to_sample`` there, so the training loop still receives the reference's normalised fp32
    ``[b, 3, h, w]`` batch, bit-identical.  Rasters must be stored at the model's size (the reference's Resize is then the
    identity); anything else raises instead of silently resampling.
    """

    def __init__(self, config):
        import glob
        self.data_list = glob.glob(config.dataset_name)
        self.config = config
        self.size = (int(config.patterns_size_height), int(config.patterns_size_width))

    def __len__(self):
        return len(self.data_list)

    def remove_sample(self, index):
        del self.data_list[index]

    def __getitem__(self, index):
        from PIL import Image
        with open(self.data_list[index], "rb") as f:
            arr = np.array(Image.open(f))
        if arr.ndim == 2:
            arr = arr[:, :, None]
        if arr.dtype != np.uint8 or arr.shape[:2] != self.size:
            raise ValueError(f"{self.data_list[index]}: expected an 8-bit raster of size {self.size}, got "
                             f"{arr.dtype} {arr.shape} (resize offline; the device path does not resample)")
        return torch.from_numpy(arr)


def is_raster_batch(t) -> bool:
    """A collated batch of RasterDataset items: uint8 [b, h, w, c] with c <= 4."""
    return torch.is_tensor(t) and t.dtype == torch.uint8 and t.dim() == 4 and 1 <= t.shape[3] <= 4
