#!/usr/bin/env python
"""Generate tests/golden/resize_golden.npz by running the REFERENCE's own ``Image_Dataset`` (with its live
``Resize((H, W), antialias=False)``, DriveSceneGen/utils/datasets/dataset.py:20-23, and its ``.pkl`` branch, :38-42) on
synthetic BEV rasters stored at sizes different from the model's.

Run in the build container (needs /root/reference, torchvision, PIL):

    python tests/golden/make_resize_golden.py

Recorded per case k: image_<k> (uint8 HWC, or float32 HWC for the .pkl case), size_<k> = (H, W), sample_<k> = what
``Image_Dataset.__getitem__`` returned, mode_<k> = which of ATen's two CPU bilinear kernels torchvision ran for that
case (0 generic / 1 channels-last; decided by comparing with the oracle's two restatements — exactly one must match),
threads = torch.get_num_threads() of the generating host.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    import torch
    from PIL import Image
    from make_raster_golden import bev_raster, load
    from oracle.raster import resize_to_sample

    dataset = load("ref_dataset", "DriveSceneGen/utils/datasets/dataset.py")
    rng = np.random.default_rng(20261018)
    #        stored (h, w, c)   model (H, W)
    cases = [((512, 512, 3), (256, 256)),    # the reference's own configuration: rasterised at 512^2, trained at 256^2
             ((400, 400, 3), (256, 256)),    # "max 400" (scripts/train.py:14): non-dyadic weights
             ((300, 200, 4), (256, 256)),    # RGBA, down in one axis and up in the other
             ((80, 52, 3), (64, 64)),        # small output: ATen takes its channels-last kernel
             ((96, 96, 3), (64, 80))]        # .pkl branch: float fig_tensor, no ToTensor
    out = {"threads": np.array(torch.get_num_threads())}
    with tempfile.TemporaryDirectory() as tmp:
        for k, ((h, w, c), (H, W)) in enumerate(cases):
            pkl = k == len(cases) - 1
            if pkl:
                img = rng.random((h, w, c), dtype=np.float32)
                path = os.path.join(tmp, f"{k}.pkl")
                torch.save({"fig_tensor": torch.from_numpy(img)}, path)
            else:
                img = bev_raster(rng, h, w, (127, 128), c)
                if k == 3:
                    img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
                path = os.path.join(tmp, f"{k}.png")
                Image.fromarray(img).save(path)
            cfg = types.SimpleNamespace(dataset_name=path, patterns_size_height=H, patterns_size_width=W)
            ds = dataset.Image_Dataset(cfg)
            assert len(ds) == 1
            sample = ds[0].numpy()
            assert sample.shape == (c, H, W) and sample.dtype == np.float32
            match = [m for m in (0, 1) if np.array_equal(resize_to_sample(img[None], H, W, mode=m)[0], sample)]
            assert len(match) >= 1, f"case {k}: neither restatement reproduces torchvision"
            out[f"image_{k}"], out[f"size_{k}"], out[f"sample_{k}"] = img, np.array([H, W]), sample
            out[f"mode_{k}"] = np.array(match[0] if len(match) == 1 else -1)   # -1: both formulas agree on this case
            print(f"case {k}: stored {h}x{w}x{c} -> {H}x{W}  modes matching torchvision: {match}")
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(HERE, "resize_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
