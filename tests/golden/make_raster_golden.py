#!/usr/bin/env python
"""Generate tests/golden/raster_golden.npz by running the REFERENCE's own functions on small synthetic BEV rasters.

Run in the build container (needs /root/reference, torchvision, cv2, PIL; matplotlib is stubbed — only used for plots):

    python tests/golden/make_raster_golden.py

What is recorded (inputs + outputs, so the fixtures travel to the GPU box without the reference):
  * Image_Dataset.__getitem__            DriveSceneGen/utils/datasets/dataset.py:15-50        -> sample_<k>
  * image_utils.get_gray_image           DriveSceneGen/vectorization/utils/image_utils.py:13  -> gray_<k>
  * extract_agents: the `thresh` image handed to cv2.findContours
                                         DriveSceneGen/vectorization/direct/extract_vehicles.py:130-148 -> agent_<k>
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def bev_raster(rng, h, w, background=(127, 128), c=3):
    """dx/dy channels: one dominant background grey + polylines of other values; speed channel: 0 + a few bright boxes."""
    img = np.zeros((h, w, c), dtype=np.uint8)
    img[..., 0] = background[0]
    img[..., 1] = background[1]
    for _ in range(6):
        y, x = rng.integers(0, h), rng.integers(0, w)
        vx, vy = int(rng.integers(0, 256)), int(rng.integers(0, 256))
        for _ in range(3 * max(h, w)):
            img[y % h, x % w, 0] = vx
            img[y % h, x % w, 1] = vy
            y += int(rng.integers(-1, 2))
            x += 1
            if rng.random() < 0.05:   # values drift along the lane, some close to the background
                vx = int(np.clip(vx + rng.integers(-20, 21), 0, 255))
                vy = int(np.clip(vy + rng.integers(-20, 21), 0, 255))
    for _ in range(5):
        y, x = rng.integers(0, h - 6), rng.integers(0, w - 10)
        img[y:y + 5, x:x + 9, 2] = rng.integers(90, 256)   # around the 100 threshold
    img[..., 2] = np.where(rng.random((h, w)) < 0.01, rng.integers(0, 256, (h, w)), img[..., 2])
    if c == 4:
        img[..., 3] = 255
    return img


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ModuleNotFoundError:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    import cv2
    import torch
    from PIL import Image
    from torchvision import transforms

    image_utils = load("ref_image_utils", "DriveSceneGen/vectorization/utils/image_utils.py")
    dataset = load("ref_dataset", "DriveSceneGen/utils/datasets/dataset.py")
    vehicles = load("ref_extract_vehicles", "DriveSceneGen/vectorization/direct/extract_vehicles.py")

    rng = np.random.default_rng(20261017)
    cases = [(64, 64, (127, 128)), (96, 128, (127, 127)), (80, 52, (0, 255)), (256, 256, (128, 127)),
             (64, 64, (255, 3))]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for k, (h, w, bg) in enumerate(cases):
            img = bev_raster(rng, h, w, bg)
            if k == 2:
                img[..., 0] = rng.integers(0, 256, (h, w))    # no dominant value: ties / low peaks
            out[f"image_{k}"] = img
            pil = Image.fromarray(img)
            pil.save(os.path.join(tmp, f"{k}.png"))

            # --- get_gray_image (the reference loops over every pixel in Python)
            gray = np.asarray(image_utils.get_gray_image(pil, plot=False))
            assert gray.shape == (h, w, 3) and (gray[..., 0] == gray[..., 1]).all() and (gray[..., 0] == gray[..., 2]).all()
            out[f"gray_{k}"] = gray[..., 0].copy()

            # --- Image_Dataset.__getitem__ at the stored size (Resize = identity)
            cfg = types.SimpleNamespace(dataset_name=os.path.join(tmp, f"{k}.png"), patterns_size_height=h,
                                        patterns_size_width=w)
            ds = dataset.Image_Dataset(cfg)
            assert len(ds) == 1
            out[f"sample_{k}"] = ds[0].numpy()

            # --- extract_agents: capture what reaches cv2.findContours, then stop (no contours -> empty list)
            seen = {}

            def fake_find_contours(image, mode, method, _seen=seen):
                _seen["thresh"] = image.copy()
                return [], None

            real = cv2.findContours
            cv2.findContours = fake_find_contours
            try:
                agents = vehicles.extract_agents(transforms.ToTensor()(pil), lanes=[])
            finally:
                cv2.findContours = real
            assert agents == []
            out[f"agent_{k}"] = seen["thresh"]
    out["n_cases"] = np.array(len(cases))
    path = os.path.join(HERE, "raster_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
