"""GPU parity of the raster-image kernels (dsg_image_to_sample, dsg_gray_mask, dsg_agent_threshold) through the C ABI:
bit-exact against the reference-generated golden vectors (tests/golden/raster_golden.npz) and against the numpy
oracle on seeded inputs, including ragged sizes, RGBA input, empty batches and full-size batches."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _rand_images(rng, n, h, w, c, dominant=True):
    img = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    if dominant:   # BEV rasters: one background value covers most of the dx / dy channels
        bg = rng.random((n, h, w)) < 0.9
        img[..., 0] = np.where(bg, 127, img[..., 0])
        img[..., 1] = np.where(bg, 128, img[..., 1])
        img[..., 2] = np.where(bg, 0, img[..., 2])
    return img


def test_golden_image_to_sample(gold):
    from drivescenegen_b200.hostapi import raster
    for k in range(int(gold["n_cases"])):
        got = raster.image_to_sample(gold[f"image_{k}"]).cpu().numpy()[0]
        assert np.array_equal(got, gold[f"sample_{k}"]), f"case {k}"


def test_golden_gray_mask(gold):
    from drivescenegen_b200.hostapi import raster
    for k in range(int(gold["n_cases"])):
        img = gold[f"image_{k}"]
        mask, peaks, hist = raster.gray_masks(img)
        assert np.array_equal(mask.cpu().numpy()[0], gold[f"gray_{k}"]), f"case {k}"
        pil = raster.get_gray_image(__import__("PIL.Image", fromlist=["Image"]).fromarray(img))
        arr = np.asarray(pil)
        assert arr.shape == img.shape[:2] + (3,)
        for ch in range(3):
            assert np.array_equal(arr[..., ch], gold[f"gray_{k}"])


def test_golden_agent_threshold(gold):
    from drivescenegen_b200.hostapi import raster
    for k in range(int(gold["n_cases"])):
        img = gold[f"image_{k}"]
        chw = torch.from_numpy(img).permute(2, 0, 1).float().div(255)      # transforms.ToTensor()
        got = raster.agent_threshold(chw.cuda()).cpu().numpy()[0]
        assert np.array_equal(got, gold[f"agent_{k}"]), f"case {k}"


@pytest.mark.parametrize("n,h,w,c", [(1, 64, 64, 3), (3, 37, 53, 3), (2, 40, 44, 4), (5, 1, 7, 3), (16, 256, 256, 3),
                                     (2, 128, 96, 4)])
@pytest.mark.parametrize("dominant", [True, False])
def test_gray_mask_vs_oracle(n, h, w, c, dominant):
    from drivescenegen_b200.hostapi import raster
    from oracle.raster import gray_mask
    rng = np.random.default_rng(n * 1000 + h + w + c)
    img = _rand_images(rng, n, h, w, c, dominant)
    gray3, peaks, hist = raster.gray_masks(img, want_gray3=True)
    mask, peaks2, _ = raster.gray_masks(img)
    gray3, peaks, hist, mask = gray3.cpu().numpy(), peaks.cpu().numpy(), hist.cpu().numpy(), mask.cpu().numpy()
    assert np.array_equal(peaks, peaks2.cpu().numpy())
    for i in range(n):
        h_ref, p_ref, m_ref = gray_mask(img[i])
        assert np.array_equal(hist[i].astype(np.int64), h_ref)
        assert np.array_equal(peaks[i], p_ref)
        assert np.array_equal(mask[i], m_ref)
        assert np.array_equal(gray3[i], np.repeat(m_ref[..., None], 3, axis=2))


def test_gray_mask_every_peak_and_threshold_edge():
    """All 256x256 (peak bin, byte value) pairs of the float64 comparison |v/255 - p/256| <= 0.1."""
    from drivescenegen_b200.hostapi import raster
    from oracle.raster import gray_mask
    imgs = []
    for p in range(0, 256):
        img = np.zeros((24, 32, 3), dtype=np.uint8)
        img[..., 0] = p              # the dx peak: 512 background pixels + one of each value
        img[..., 1] = 255 - p
        img[:8, :, 0] = np.arange(256, dtype=np.uint8).reshape(8, 32)
        img[8:16, :, 1] = np.arange(256, dtype=np.uint8).reshape(8, 32)
        imgs.append(img)
    imgs = np.stack(imgs)
    mask, peaks, _ = raster.gray_masks(imgs)
    mask, peaks = mask.cpu().numpy(), peaks.cpu().numpy()
    for p in range(256):
        _, p_ref, m_ref = gray_mask(imgs[p])
        assert np.array_equal(peaks[p], p_ref)
        assert np.array_equal(mask[p], m_ref), f"peak {p}"


@pytest.mark.parametrize("n,h,w,c,co", [(1, 64, 64, 3, 3), (3, 37, 53, 3, 3), (2, 40, 44, 4, 3), (32, 256, 256, 3, 3),
                                        (2, 16, 16, 1, 1), (2, 16, 18, 2, 1)])
def test_image_to_sample_vs_oracle(n, h, w, c, co):
    from drivescenegen_b200.hostapi import raster
    from oracle.raster import image_to_sample
    rng = np.random.default_rng(7 + n + h)
    img = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    got = raster.image_to_sample(img, channels=co).cpu().numpy()
    assert np.array_equal(got, image_to_sample(img, co))


def test_image_to_sample_all_byte_values_match_torchvision_arithmetic():
    from drivescenegen_b200.hostapi import raster
    v = torch.arange(256, dtype=torch.uint8).reshape(1, 16, 16, 1).repeat(1, 1, 1, 3)
    got = raster.image_to_sample(v).cpu()
    ref = v.permute(0, 3, 1, 2).to(torch.float32).div(255).sub(0.5).div(0.5)
    assert torch.equal(got, ref)
    assert got.min() == -1.0 and got.max() == 1.0


@pytest.mark.parametrize("n,h,w", [(1, 64, 64), (3, 37, 53), (16, 256, 256)])
def test_agent_threshold_vs_oracle(n, h, w):
    from drivescenegen_b200.hostapi import raster
    from oracle.raster import agent_threshold
    rng = np.random.default_rng(11 + n)
    u8 = rng.integers(0, 256, (n, 3, h, w), dtype=np.uint8)
    x = torch.from_numpy(u8).float().div(255)
    got = raster.agent_threshold(x.cuda()).cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], agent_threshold(x[i, 2].numpy()))
    # all byte values: the truncation of (v/255)*255 matters around the threshold
    allv = torch.arange(256, dtype=torch.float32).div(255).reshape(1, 1, 16, 16).repeat(1, 3, 1, 1)
    got = raster.agent_threshold(allv.cuda()).cpu().numpy()[0]
    assert np.array_equal(got, agent_threshold(allv[0, 2].numpy()))


def lib_err(call):
    """the keyword a failing C-ABI call leaves in dsg_last_error()."""
    from drivescenegen_b200 import _lib
    lib = _lib.load()
    assert call(lib) != 0
    msg = lib.dsg_last_error()
    for key in (b"null pointer", b"mode", b"bad shape"):
        if key in msg:
            return key
    return msg


def test_raster_edge_cases_and_errors():
    from drivescenegen_b200 import _lib
    from drivescenegen_b200.hostapi import raster
    dev = torch.device("cuda", 0)
    empty = torch.empty((0, 8, 8, 3), dtype=torch.uint8, device=dev)
    assert raster.image_to_sample(empty).shape == (0, 3, 8, 8)
    m, p, h = raster.gray_masks(empty)
    assert m.shape == (0, 8, 8) and p.shape == (0, 3)
    # zero pixels: numpy's argmax of an all-zero histogram is bin 0
    m, p, h = raster.gray_masks(torch.empty((2, 0, 8, 3), dtype=torch.uint8, device=dev))
    assert p.cpu().tolist() == [[0, 0, 0], [0, 0, 0]] and int(h.sum()) == 0
    with pytest.raises(ValueError):
        raster.gray_masks(torch.zeros((1, 8, 8, 2), dtype=torch.uint8, device=dev))
    with pytest.raises(ValueError):
        raster.image_to_sample(torch.zeros((1, 8, 8, 3), dtype=torch.float64, device=dev))
    # float32 rasters are the .pkl branch (already in [0, 1]): Normalize only
    f = torch.rand((1, 8, 8, 3), device=dev)
    assert torch.equal(raster.image_to_sample(f), f.permute(0, 3, 1, 2).sub(0.5).div(0.5))
    assert lib_err(lambda l: l.dsg_resize_to_sample(None, 0, None, 1, 8, 8, 3, 3, 4, 4, 0, None)) == b"null pointer"
    assert lib_err(lambda l: l.dsg_resize_to_sample(None, 0, None, 1, 8, 8, 3, 3, 4, 4, 2, None)) == b"mode"
    with pytest.raises(ValueError):
        raster.agent_threshold(torch.zeros((3, 8, 8), dtype=torch.float64, device=dev))
    with pytest.raises(ValueError):
        raster.get_gray_image(np.zeros((8, 8, 3), np.uint8), plot=True)
    lib = _lib.load()
    assert lib.dsg_gray_mask(None, None, None, None, None, 1, 8, 8, 3, 0.1, None) != 0
    assert b"null pointer" in lib.dsg_last_error()
    assert lib.dsg_image_to_sample(None, None, 1, 8, 8, 5, 3, None) != 0


def test_generated_sample_to_vector_front_end_stays_on_device():
    """latent -> uint8 raster (dsg_latent_to_image) -> grey mask + agent blobs, all on the device, equals the reference
    chain numpy_to_pil -> get_gray_image / ToTensor -> extract_agents head (oracle)."""
    from drivescenegen_b200 import ops
    from drivescenegen_b200.hostapi import raster
    from oracle.raster import agent_threshold, gray_mask
    g = torch.Generator().manual_seed(3)
    lat = torch.randn((4, 3, 64, 64), generator=g).mul(0.2)
    lat[:, 0] += 0.0
    lat[:, 2] -= 0.6
    u8, _ = ops.latent_to_image(lat.cuda(), want_u8=True, want_f32=False)
    mask, _, _ = raster.gray_masks(u8)
    blobs = raster.agent_threshold(u8.permute(0, 3, 1, 2).float().div(255))
    ref_u8 = (lat / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
    ref_u8 = (ref_u8 * 255).round().astype(np.uint8)
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    for i in range(4):
        assert np.array_equal(mask[i].cpu().numpy(), gray_mask(ref_u8[i])[2])
        plane = torch.from_numpy(ref_u8[i]).permute(2, 0, 1).float().div(255)[2].numpy()
        assert np.array_equal(blobs[i].cpu().numpy(), agent_threshold(plane))


def test_raster_dataset_through_accelerator_equals_reference_dataset_arithmetic(tmp_path):
    """RasterDataset -> DataLoader -> Accelerator.prepare: uint8 bytes cross PCIe, the device kernel normalises; the
    batch the training loop sees equals ToTensor + Normalize([0.5],[0.5]) (the golden test pins that to Image_Dataset)."""
    import types
    from PIL import Image
    from drivescenegen_b200.hostapi import Accelerator, RasterDataset
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (6, 32, 48, 3), dtype=np.uint8)
    for i, im in enumerate(imgs):
        Image.fromarray(im).save(tmp_path / f"{i:02d}.png")
    cfg = types.SimpleNamespace(dataset_name=str(tmp_path / "*.png"), patterns_size_height=32, patterns_size_width=48)
    ds = RasterDataset(cfg)
    ds.data_list.sort()
    assert len(ds) == 6 and ds[0].dtype == torch.uint8 and ds[0].shape == (32, 48, 3)
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False)
    acc = Accelerator()
    loader = acc.prepare(loader)
    batches = list(loader)
    assert [tuple(b.shape) for b in batches] == [(4, 3, 32, 48), (2, 3, 32, 48)]
    got = torch.cat(batches).cpu()
    ref = torch.from_numpy(imgs).permute(0, 3, 1, 2).float().div(255).sub(0.5).div(0.5)
    assert got.is_floating_point() and torch.equal(got, ref)


RESIZE_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_golden.npz")


def test_golden_resize_to_sample():
    """dsg_resize_to_sample against what the reference's Image_Dataset (live Resize, .pkl branch) returned."""
    from drivescenegen_b200.hostapi import raster
    g = np.load(RESIZE_GOLD)
    for k in range(int(g["n_cases"])):
        img, (H, W), ref, mode = g[f"image_{k}"], g[f"size_{k}"], g[f"sample_{k}"], int(g[f"mode_{k}"])
        src = torch.from_numpy(img) if img.dtype == np.float32 else img
        got = raster.image_to_sample(src, channels=img.shape[2], size=(int(H), int(W)), mode=mode).cpu().numpy()[0]
        assert np.array_equal(got, ref), f"case {k} (explicit mode)"
        # the default rule (host thread count of THIS box; the fixtures were made on a multi-core host)
        if torch.get_num_threads() > 1:
            auto = raster.image_to_sample(src, channels=img.shape[2], size=(int(H), int(W))).cpu().numpy()[0]
            assert np.array_equal(auto, ref), f"case {k} (automatic mode)"


@pytest.mark.parametrize("n,h,w,c,co,H,W", [(2, 512, 512, 3, 3, 256, 256), (3, 37, 53, 4, 3, 64, 96),
                                            (1, 16, 16, 3, 2, 64, 64), (2, 100, 60, 1, 1, 60, 100),
                                            (16, 512, 512, 3, 3, 256, 256)])
@pytest.mark.parametrize("mode", [0, 1])
def test_resize_to_sample_vs_oracle(n, h, w, c, co, H, W, mode):
    from drivescenegen_b200.hostapi import raster
    from oracle.raster import resize_to_sample
    rng = np.random.default_rng(h * 7 + w)
    img = _rand_images(rng, n, h, w, c, dominant=(c >= 3))
    check = slice(0, 2)   # the numpy oracle on two images is enough at the big batch
    got = raster.image_to_sample(img, channels=co, size=(H, W), mode=mode).cpu().numpy()
    assert got.shape == (n, co, H, W)
    assert np.array_equal(got[check], resize_to_sample(img[check], H, W, c_out=co, mode=mode))
    if n > 2:   # size-independent property: every image of the batch is processed alike
        again = raster.image_to_sample(img[-1:], channels=co, size=(H, W), mode=mode).cpu().numpy()
        assert np.array_equal(got[-1:], again)
    f = torch.from_numpy(img[check].astype(np.float32) / np.float32(255.0))
    gotf = raster.image_to_sample(f, channels=co, size=(H, W), mode=mode).cpu().numpy()
    assert np.array_equal(gotf, resize_to_sample(f.numpy(), H, W, c_out=co, mode=mode))


def test_raster_dataset_with_stored_512_rasters_feeds_256_batches(tmp_path):
    """The reference's real data layout: rasters stored at 512^2 (config/data_rasterization.yaml), model at 256^2
    (scripts/train.py:14-15).  RasterDataset -> DataLoader -> Accelerator.prepare must hand the loop exactly what
    Image_Dataset would have: checked against the reference-generated golden sample of case 0."""
    import types
    from PIL import Image
    from drivescenegen_b200.hostapi import Accelerator, RasterDataset
    g = np.load(RESIZE_GOLD)
    img, ref = g["image_0"], g["sample_0"]
    for i in range(3):
        Image.fromarray(img).save(tmp_path / f"{i}.png")
    cfg = types.SimpleNamespace(dataset_name=str(tmp_path / "*.png"), patterns_size_height=256, patterns_size_width=256)
    loader = Accelerator().prepare(torch.utils.data.DataLoader(RasterDataset(cfg), batch_size=2, shuffle=False))
    batches = list(loader)
    assert [tuple(b.shape) for b in batches] == [(2, 3, 256, 256), (1, 3, 256, 256)]
    for b in batches:
        for smp in b.cpu().numpy():
            assert np.array_equal(smp, ref)
    # .pkl branch end to end (case 4 of the fixture)
    k = int(g["n_cases"]) - 1
    fig, (H, W), refp = g[f"image_{k}"], g[f"size_{k}"], g[f"sample_{k}"]
    torch.save({"fig_tensor": torch.from_numpy(fig)}, tmp_path / "x.pkl")
    cfgp = types.SimpleNamespace(dataset_name=str(tmp_path / "*.pkl"), patterns_size_height=int(H),
                                 patterns_size_width=int(W))
    (bp,) = list(Accelerator().prepare(torch.utils.data.DataLoader(RasterDataset(cfgp), batch_size=1)))
    assert np.array_equal(bp.cpu().numpy()[0], refp)
