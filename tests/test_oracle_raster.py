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
    assert raster.is_raster_batch(batch.float(), ds)                      # float rasters only when the dataset says so
    assert not raster.is_raster_batch(torch.zeros(2, 3, 8, 8))            # a normalised fp32 batch is left alone
    # stored size != model size is fine (the device resamples); the .pkl branch returns the float fig_tensor, and a
    # non-dict pickle falls through to the next item (dataset.py:38-42)
    other = types.SimpleNamespace(dataset_name=str(tmp_path / "*.png"), patterns_size_height=32, patterns_size_width=32)
    assert tuple(RasterDataset(other)[0].shape) == (24, 40, 3) and RasterDataset(other).size == (32, 32)
    fig = torch.rand(10, 12, 3)
    torch.save([1, 2, 3], tmp_path / "a.pkl")
    torch.save({"fig_tensor": fig}, tmp_path / "b.pkl")
    pk = RasterDataset(types.SimpleNamespace(dataset_name=str(tmp_path / "*.pkl"), patterns_size_height=8,
                                             patterns_size_width=8))
    pk.data_list.sort()
    assert torch.equal(pk[0], fig) and torch.equal(pk[1], fig)
    if not torch.cuda.is_available():
        for call in (lambda: raster.image_to_sample(imgs), lambda: raster.gray_masks(imgs),
                     lambda: raster.get_gray_image(Image.fromarray(imgs[0])),
                     lambda: raster.agent_threshold(torch.zeros(3, 8, 8))):
            with pytest.raises(DsgError):
                call()
    with pytest.raises(ValueError):
        raster.image_to_sample(imgs.astype(np.float64))


RESIZE_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_golden.npz")


def test_resize_oracle_matches_reference_dataset_with_live_resize():
    """Image_Dataset with stored size != model size (the reference's 512^2 -> 256^2, non-dyadic 400 -> 256, RGBA, a small
    output, the .pkl branch): the oracle's restatement of ATen's two CPU bilinear kernels reproduces torchvision bit for
    bit, and the kernel-selection rule the host API applies picks the variant torchvision actually ran."""
    from drivescenegen_b200.hostapi.raster import aten_resize_mode
    from oracle.raster import resize_to_sample
    g = np.load(RESIZE_GOLD)
    assert int(g["n_cases"]) >= 5 and int(g["threads"]) > 1
    seen = set()
    for k in range(int(g["n_cases"])):
        img, (H, W), ref, mode = g[f"image_{k}"], g[f"size_{k}"], g[f"sample_{k}"], int(g[f"mode_{k}"])
        assert mode in (0, 1), "every fixture case distinguishes the two kernels"
        got = resize_to_sample(img[None], int(H), int(W), mode=mode)[0]
        assert got.dtype == np.float32 and np.array_equal(got, ref), f"case {k}"
        assert not np.array_equal(resize_to_sample(img[None], int(H), int(W), mode=1 - mode)[0], ref)
        assert aten_resize_mode(int(H), int(W), img.shape[2], host_threads=int(g["threads"])) == mode
        seen.add((mode, img.dtype.name))
    assert {(0, "uint8"), (1, "uint8"), (0, "float32")} <= seen


def test_resize_oracle_identity_and_fma():
    from oracle.raster import fma32, image_to_sample, resize_to_sample
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (2, 20, 28, 3), dtype=np.uint8)
    assert np.array_equal(resize_to_sample(img, 20, 28), image_to_sample(img))      # same size: Resize is a copy
    # fma32 is the exactly rounded fused multiply-add: against exact rational arithmetic on awkward operands
    from fractions import Fraction
    a = rng.standard_normal(2000).astype(np.float32)
    b = rng.standard_normal(2000).astype(np.float32)
    c = (-(a.astype(np.float64) * b.astype(np.float64))).astype(np.float32) + rng.standard_normal(2000).astype(np.float32) * 1e-7
    got = fma32(a, b, c)
    for i in range(0, 2000, 7):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        lo, hi = np.nextafter(got[i], np.float32(-np.inf)), np.nextafter(got[i], np.float32(np.inf))
        assert abs(Fraction(float(got[i])) - exact) <= min(abs(Fraction(float(lo)) - exact), abs(Fraction(float(hi)) - exact))
