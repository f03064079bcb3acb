"""GPU checks of the remaining BASELINE.json configs: 512x512 DDIM sampling (configs[3]) through the pipeline API and
size-independent properties at full size."""
import pytest
import torch

pytestmark = pytest.mark.gpu

REF_CFG = dict(in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(64, 128, 256, 512),
               down_block_types=("DownBlock2D",) * 4, up_block_types=("UpBlock2D",) * 4)


def _dev():
    return torch.device("cuda", 0)


def test_unet_512_matches_oracle():
    """one 512x512 forward (configs[3] resolution; mid-block attention over 4096 tokens) vs the fp32 CPU oracle.
    Tolerance: relative L2 <= 1e-2 (fp16 operands, fp32 accumulation)."""
    from drivescenegen_b200.hostapi import UNet2DModel
    from oracle.unet import OracleUNet2D
    torch.manual_seed(0)
    oracle = OracleUNet2D(sample_size=512, **REF_CFG).eval()
    model = UNet2DModel(sample_size=512, **REF_CFG)
    model.load_state_dict(oracle.state_dict())
    model = model.to(_dev()).eval()
    x = torch.randn(1, 3, 512, 512, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = oracle(x, 321)[0]
        got = model(x.to(_dev()), 321).sample.cpu()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 1e-2, rel


def test_ddim_pipeline_50_steps_512_properties():
    """configs[3]: 50-step DDIM sampling at 512x512 through DDIMPipeline (CUDA-graph path).  Properties that do not need
    the CPU oracle at this size: shape/range of the images, determinism for a fixed seed, batch independence of one
    denoising step (sample 0 of a batch-2 forward equals the batch-1 forward: samples never interact; a 50-step
    trajectory of a random-init U-Net is chaotic, so the comparison is per step)."""
    from drivescenegen_b200.hostapi import DDIMPipeline, DDIMScheduler, UNet2DModel
    torch.manual_seed(0)
    model = UNet2DModel(sample_size=512, **REF_CFG).to(_dev()).eval()
    pipe = DDIMPipeline(model, DDIMScheduler())
    pipe.set_progress_bar_config(disable=True)
    a = pipe(batch_size=2, generator=torch.manual_seed(3), num_inference_steps=50, output_type="np").images
    b = pipe(batch_size=2, generator=torch.manual_seed(3), num_inference_steps=50, output_type="np").images
    assert a.shape == (2, 512, 512, 3) and a.dtype.name == "float32"
    assert a.min() >= 0.0 and a.max() <= 1.0
    assert (a == b).all()
    x = torch.randn(2, 3, 512, 512, generator=torch.Generator().manual_seed(4)).to(_dev())
    with torch.no_grad():
        e2 = model(x, 980).sample
        e1 = model(x[:1].contiguous(), 980).sample
    # per-sample arithmetic is the same up to the chunking of conv_in's GroupNorm statistics (fp32 partials, ~1e-7),
    # which can flip a few fp16 roundings: agreement at fp16 resolution, not bit-identity
    assert (e2[:1] - e1).abs().max().item() <= 5e-3 * e1.abs().max().item()
    assert ((e2[:1] - e1).norm() / e1.norm()).item() < 1e-3
