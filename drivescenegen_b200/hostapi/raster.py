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


def aten_resize_mode(out_h: int, out_w: int, channels: int, host_threads: int = None) -> int:
    """Which of ATen's two CPU bilinear kernels the reference's ``Resize(antialias=False)`` would have run on the host
    (they differ in the last bit): 1 = the channels-last kernel, taken by single-threaded hosts for 3-channel images and
    (observed with torch 2.11, recorded in tests/golden/resize_golden.npz) by outputs with ``out_h + out_w <= 128``;
    0 = the generic N-d kernel, i.e. the reference's 512^2 -> 256^2 on any multi-core host."""
    if host_threads is None:
        host_threads = torch.get_num_threads()
    if out_h + out_w <= 128 or (host_threads == 1 and channels == 3):
        return 1
    return 0


def image_to_sample(images, channels: int = 3, out: torch.Tensor = None, device=None, size=None,
                    mode: int = None) -> torch.Tensor:
    """rasters [n,h,w,c] (or one [h,w,c]) -> normalised fp32 samples [n,channels,H,W] in [-1, 1].

    uint8 rasters go through ToTensor (x / 255); float32 rasters (the ``.pkl`` branch of ``Image_Dataset``) are taken as
    they are.  ``size=(H, W)`` different from the stored size adds the transform's ``Resize(size, antialias=False)``
    between ToTensor and Normalize, bit-identical to torchvision on the host; ``mode`` picks ATen's kernel variant
    (default: :func:`aten_resize_mode`)."""
    if isinstance(images, torch.Tensor) and images.dtype == torch.float32:
        img = images[None] if images.dim() == 3 else images
        if img.dim() != 4:
            raise ValueError(f"expected [h,w,c] or [n,h,w,c] rasters, got shape {tuple(images.shape)}")
        if not img.is_cuda:
            img = img.to(_device(device), non_blocking=True)
        img = img.contiguous()
    else:
        img = _as_u8_batch(images, device)
    n, h, w, c = img.shape
    if not 1 <= channels <= c:
        raise ValueError(f"channels={channels} but the images have {c}")
    oh, ow = (h, w) if size is None else (int(size[0]), int(size[1]))
    if out is None:
        out = torch.empty((n, channels, oh, ow), dtype=torch.float32, device=img.device)
    elif out.shape != (n, channels, oh, ow) or out.dtype != torch.float32 or not out.is_contiguous() \
            or out.device != img.device:
        raise ValueError("out must be a contiguous fp32 tensor [n,channels,H,W] on the images' device")
    lib = _lib.load()
    with torch.cuda.device(img.device):
        if (oh, ow) == (h, w) and img.dtype == torch.uint8 and c <= 4:
            check(lib.dsg_image_to_sample(img.data_ptr(), out.data_ptr(), n, h, w, c, channels, _stream(img.device)),
                  "dsg_image_to_sample")
        else:
            if mode is None:
                mode = aten_resize_mode(oh, ow, channels)
            check(lib.dsg_resize_to_sample(img.data_ptr(), 1 if img.dtype == torch.float32 else 0, out.data_ptr(), n, h,
                                           w, c, channels, oh, ow, int(mode), _stream(img.device)),
                  "dsg_resize_to_sample")
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
    with torch.cuda.device(dev):
        check(_lib.load().dsg_gray_mask(img.data_ptr(), hist.data_ptr(), peaks.data_ptr(), mask.data_ptr(),
                                        gray3.data_ptr() if want_gray3 else None, n, h, w, c, float(thresh),
                                        _stream(dev)), "dsg_gray_mask")
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
    with torch.cuda.device(x.device):
        check(_lib.load().dsg_agent_threshold(plane, c * h * w, out.data_ptr(), n, h * w, int(thresh),
                                              _stream(x.device)), "dsg_agent_threshold")
    return out


class RasterDataset(torch.utils.data.Dataset):
    """``Image_Dataset`` (DriveSceneGen/utils/datasets/dataset.py:15-50) with the arithmetic left for the device.

    Same constructor argument (``config.dataset_name`` glob, ``patterns_size_height/width``), same ``data_list`` /
    ``remove_sample``; ``__getitem__`` only decodes the file and returns the raw ``[h, w, c]`` raster at its STORED size:
    uint8 for images, float32 for the ``.pkl`` branch (``torch.load(f)['fig_tensor']``, dataset.py:38-42; a non-dict
    pickle falls through to the next item like the reference).  A DataLoader over it collates ``[b, h, w, c]`` batches;
    ``Accelerator.prepare`` (ShardedDataLoader) ships those bytes to the GPU and runs ToTensor + ``Resize((H, W),
    antialias=False)`` + ``Normalize`` there (``dsg_image_to_sample`` / ``dsg_resize_to_sample``), so the training loop
    still receives the reference's fp32 ``[b, 3, H, W]`` batch, bit-identical.  All files of one dataset must share one
    stored size (the reference's rasteriser writes 512^2) — the batch is collated before it is resampled.
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
        import os
        file = self.data_list[index]
        if os.path.splitext(file)[1].lower() == ".pkl":
            with open(file, "rb") as f:
                data = torch.load(f, weights_only=False)
            if not isinstance(data, dict):
                return self.__getitem__(index + 1)
            fig = data["fig_tensor"][:, :, :]
            if fig.dtype != torch.float32:
                raise ValueError(f"{file}: fig_tensor must be float32, got {fig.dtype}")
            return fig.contiguous()
        from PIL import Image
        with open(file, "rb") as f:
            arr = np.array(Image.open(f))
        if arr.ndim == 2:
            arr = arr[:, :, None]
        if arr.dtype != np.uint8:
            raise ValueError(f"{file}: expected an 8-bit raster, got {arr.dtype}")
        return torch.from_numpy(arr)


def is_raster_batch(t, dataset=None) -> bool:
    """A collated batch of RasterDataset items: [b, h, w, c] uint8 with c <= 4 (or any dtype RasterDataset returns when
    the loader's dataset is known to be one)."""
    if not torch.is_tensor(t) or t.dim() != 4:
        return False
    if isinstance(dataset, RasterDataset):
        return t.dtype in (torch.uint8, torch.float32)
    return t.dtype == torch.uint8 and 1 <= t.shape[3] <= 4
