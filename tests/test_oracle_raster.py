"""The raster oracle (oracle/raster.py) against golden vectors produced by the REFERENCE's own functions
(tests/golden/make_raster_golden.py -> raster_golden.npz): Image_Dataset.__getitem__, get_gray_image, and the threshold
image extract_agents hands to cv2.findContours.  Bit-exact (byte / exactly-rounded fp32 work)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def cases(g):
    return range(int(g["n_cases"]))


def test_golden_file_has_every_case(gold):
    assert int(gold["n_cases"]) >= 5
    for k in cases(gold):
        for name in ("image", "gray", "sample", "agent"):
            assert f"{name}_{k}" in gold.files


def test_image_to_sample_matches_reference_dataset(gold):
    from oracle.raster import image_to_sample
    for k in cases(gold):
        got = image_to_sample(gold[f"image_{k}"][None])[0]
        ref = gold[f"sample_{k}"]
        assert got.dtype == np.float32 and got.shape == ref.shape
        assert np.array_equal(got, ref), f"case {k}"


def test_gray_mask_matches_reference_get_gray_image(gold):
    from oracle.raster import gray_mask
    for k in cases(gold):
        img = gold[f"image_{k}"]
        hist, peaks, mask = gray_mask(img)
        assert np.array_equal(mask, gold[f"gray_{k}"]), f"case {k}"
        assert hist.sum(axis=1).tolist() == [img.shape[0] * img.shape[1]] * 3
        for ch in range(3):   # every byte value lands in bin min(v, 255); 255 shares the last bin with nothing else
            assert np.array_equal(hist[ch], np.bincount(img[..., ch].ravel(), minlength=256))
            assert peaks[ch] == int(np.argmax(hist[ch]))


def test_agent_threshold_matches_reference_extract_agents(gold):
    from oracle.raster import agent_threshold
    for k in cases(gold):
        img = gold[f"image_{k}"]
        plane = img[..., 2].astype(np.float32) / np.float32(255.0)     # transforms.ToTensor()
        assert np.array_equal(agent_threshold(plane), gold[f"agent_{k}"]), f"case {k}"


def test_agent_threshold_truncation_cases():
    """(v/255)*255 in fp32 falls just below v for some v, and astype(uint8) truncates: the blob mask is NOT `v > 100`."""
    from oracle.raster import agent_threshold
    v = np.arange(256, dtype=np.uint8)
    plane = (v.astype(np.float32) / np.float32(255.0)).reshape(16, 16)
    trunc = (plane * 255).astype(np.uint8)
    got = agent_threshold(plane).ravel()
    assert np.array_equal(got, np.where(trunc.ravel() > 100, 255, 0))
    assert got[100] == 0 and got[102] == 255


def test_raster_dataset_host_side_and_no_cpu_path(tmp_path):
    """RasterDataset (Image_Dataset's constructor / data_list / remove_sample; decode only) and the rule that the raster
    kernels have no CPU fallback: without a CUDA device every entry point raises DsgError."""
    import types

    import torch
    from PIL import Image

    from drivescenegen_b200._lib import DsgError
    from drivescenegen_b200.hostapi import RasterDataset, raster
    rng = np.random.default_rng(1)
    imgs = rng.integers(0, 256, (3, 24, 40, 3), dtype=np.uint8)
    for i, im in enumerate(imgs):
        Image.fromarray(im).save(tmp_path / f"{i}.png")
    cfg = types.SimpleNamespace(dataset_name=str(tmp_path / "*.png"), patterns_size_height=24, patterns_size_width=40)
    ds = RasterDataset(cfg)
    ds.data_list.sort()
    assert len(ds) == 3
    item = ds[1]
    assert item.dtype == torch.uint8 and tuple(item.shape) == (24, 40, 3) and np.array_equal(item.numpy(), imgs[1])
    ds.remove_sample(0)
    assert len(ds) == 2 and np.array_equal(ds[0].numpy(), imgs[1])
    batch = torch.utils.data.default_collate([ds[0], ds[1]])
    assert raster.is_raster_batch(batch) and not raster.is_raster_batch(batch.float())
    assert not raster.is_raster_batch(torch.zeros(2, 3, 8, 8))            # a normalised fp32 batch is left alone
    wrong = types.SimpleNamespace(dataset_name=str(tmp_path / "*.png"), patterns_size_height=32, patterns_size_width=32)
    with pytest.raises(ValueError):
        RasterDataset(wrong)[0]
    if not torch.cuda.is_available():
        for call in (lambda: raster.image_to_sample(imgs), lambda: raster.gray_masks(imgs),
                     lambda: raster.get_gray_image(Image.fromarray(imgs[0])),
                     lambda: raster.agent_threshold(torch.zeros(3, 8, 8))):
            with pytest.raises(DsgError):
                call()
    with pytest.raises(ValueError):
        raster.image_to_sample(imgs.astype(np.float32))
