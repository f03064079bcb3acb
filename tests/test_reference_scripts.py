"""CPU plumbing (BASELINE.json configs[0]): the reference's UNMODIFIED `DriveSceneGen/pipeline/training_pipeline.py`
and `DriveSceneGen/utils/datasets/dataset.py` run against the `diffusers` / `accelerate` shims.

There is no GPU here, so the test-suite monkeypatches the host-API classes to serve CPU tensors from the CPU oracle
(`tests/cpu_plumbing.py`, test infrastructure; the product has no CPU path and no hook for one).  Skipped where
/root/reference is absent (the GPU box) — the GPU box runs the staged copies in `tests/test_gpu_reference_scripts.py`."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "DriveSceneGen")),
                                reason="reference checkout not present")


@pytest.fixture()
def oracle_backend(monkeypatch):
    import cpu_plumbing
    cpu_plumbing.install(monkeypatch)
    yield


def test_training_pipeline_runs_unmodified_on_cpu(tmp_path, monkeypatch, oracle_backend):
    from PIL import Image
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # the package __init__ configures logging from a yaml next to it; import the two modules we need directly
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    tp = load("ref_training_pipeline", "DriveSceneGen/pipeline/training_pipeline.py")
    ds = load("ref_dataset", "DriveSceneGen/utils/datasets/dataset.py")
    from diffusers import DDPMPipeline, DDPMScheduler, UNet2DModel
    from diffusers.optimization import get_cosine_schedule_with_warmup
    from accelerate import notebook_launcher
    assert tp.DDPMPipeline is DDPMPipeline  # the reference module really imported the shim

    data = tmp_path / "data"
    data.mkdir()
    rng = np.random.default_rng(0)
    for i in range(4):
        Image.fromarray(rng.integers(0, 255, (80, 80, 3), dtype=np.uint8)).save(data / f"{i}.png")

    class Cfg:  # BASELINE configs[0]: 64x64, 2 blocks (the reference's TrainingConfig fields, scripts/train.py:12-28)
        patterns_size_height = 64
        patterns_size_width = 64
        train_batch_size = 2
        eval_batch_size = 1
        num_epochs = 1
        gradient_accumulation_steps = 1
        learning_rate = 1e-4
        lr_warmup_steps = 1
        save_image_epochs = 1
        save_model_epochs = 1
        mixed_precision = "no"
        output_dir = str(tmp_path / "out")
        dataset_name = str(data / "*")
        seed = 14555

    cfg = Cfg()
    dataset = ds.Image_Dataset(cfg)
    loader = torch.utils.data.DataLoader(dataset, batch_size=cfg.train_batch_size, shuffle=True)
    torch.manual_seed(0)
    model = UNet2DModel(sample_size=(64, 64), in_channels=3, out_channels=3, layers_per_block=2,
                        block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2,
                        up_block_types=("UpBlock2D",) * 2)
    w0 = model.conv_out.weight.detach().clone()
    sched = DDPMScheduler()
    opt = torch.optim.AdamW(model.parameters(), lr=cfg.learning_rate)
    lrs = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=cfg.lr_warmup_steps,
                                          num_training_steps=len(loader) * cfg.num_epochs)
    pipeline = tp.TrainingPipeline(cfg)
    # evaluate() hard-codes 750 sampling steps (training_pipeline.py:26-32); trim the scheduler's work for CI by
    # monkeypatching nothing in the reference: 750 steps of the 3.7M-param model at 64x64 take ~15 s on 8 threads
    notebook_launcher(pipeline.train_loop, (cfg, model, sched, opt, loader, lrs), num_processes=1)
    assert not torch.equal(model.conv_out.weight.detach(), w0), "optimizer did not update the weights"
    out = tmp_path / "out"
    assert (out / "model_index.json").is_file() and (out / "unet" / "diffusion_pytorch_model.bin").is_file()
    samples = sorted(os.listdir(out / "samples"))
    assert samples == ["000.png"]
    assert Image.open(out / "samples" / "000.png").size == (64, 64)
    assert any(f.startswith("events.out.tfevents") for f in os.listdir(out / "logs" / "train_example"))
    # checkpoint loads the way scripts/generation.py:7 and scripts/train.py:59 do
    p2 = DDPMPipeline.from_pretrained(str(out), variant="fp16")
    assert torch.equal(p2.unet.conv_out.weight, model.conv_out.weight.detach().cpu())
    m2 = UNet2DModel.from_pretrained(str(out), subfolder="unet")
    assert m2.config.block_out_channels == [64, 128] or tuple(m2.config.block_out_channels) == (64, 128)
