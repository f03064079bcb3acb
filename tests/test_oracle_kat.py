"""Pins for the CPU oracle: upstream scheduler known-answer loops and parameter counts."""
import torch

from oracle.schedulers import OracleDDIMScheduler, OracleDDPMScheduler
from oracle.unet import OracleUNet2D


def _dummy_sample_deter():
    n = 4 * 3 * 8 * 8
    s = torch.arange(n).reshape(3, 8, 8, 4) / n
    return s.permute(3, 0, 1, 2)


def _dummy_model(sample, t):
    return sample * t / (t + 1)


def test_ddpm_full_loop_no_noise_kat():
    # upstream tests/schedulers/test_scheduler_ddpm.py::test_full_loop_no_noise -> 258.9606 / 0.3372
    sch = OracleDDPMScheduler()
    sample = _dummy_sample_deter()
    gen = torch.manual_seed(0)
    for t in reversed(range(len(sch))):
        sample = sch.step(_dummy_model(sample, t), t, sample, generator=gen)
    assert abs(sample.abs().sum().item() - 258.9606) < 1e-2
    assert abs(sample.abs().mean().item() - 0.3372) < 1e-3


def test_ddim_full_loop_no_noise_kat():
    # upstream tests/schedulers/test_scheduler_ddim.py::test_full_loop_no_noise -> 172.0067 / 0.223967
    sch = OracleDDIMScheduler()
    sch.set_timesteps(10)
    sample = _dummy_sample_deter()
    for t in sch.timesteps:
        sample = sch.step(_dummy_model(sample, t), t, sample, eta=0.0)
    assert abs(sample.abs().sum().item() - 172.0067) < 1e-2
    assert abs(sample.abs().mean().item() - 0.223967) < 1e-3


def test_timestep_spacing_leading():
    sch = OracleDDPMScheduler()
    sch.set_timesteps(750)
    assert sch.timesteps[0].item() == 749 and sch.timesteps[-1].item() == 0 and len(sch.timesteps) == 750
    sch.set_timesteps(50)
    assert sch.timesteps[0].item() == 980 and sch.timesteps[1].item() == 960


def _count(m):
    return sum(p.numel() for p in m.parameters())


def test_param_count_reference_config():
    # DriveSceneGen/scripts/train.py:39-57 -> 56,574,595
    m = OracleUNet2D(sample_size=(256, 256), in_channels=3, out_channels=3, layers_per_block=2,
                     block_out_channels=(64, 128, 256, 512), down_block_types=("DownBlock2D",) * 4,
                     up_block_types=("UpBlock2D",) * 4)
    assert _count(m) == 56_574_595
    keys = set(m.state_dict().keys())
    for k in ["conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.1.resnets.0.conv_shortcut.weight",
              "down_blocks.2.downsamplers.0.conv.bias", "mid_block.attentions.0.to_out.0.weight",
              "mid_block.attentions.0.group_norm.weight", "up_blocks.0.resnets.2.conv_shortcut.bias",
              "up_blocks.2.upsamplers.0.conv.weight", "conv_norm_out.weight", "conv_out.bias"]:
        assert k in keys, k
    assert "down_blocks.0.resnets.0.conv_shortcut.weight" not in keys
    assert "down_blocks.3.downsamplers.0.conv.weight" not in keys
    assert "up_blocks.3.upsamplers.0.conv.weight" not in keys


def test_param_count_hf_tutorial_config():
    # public HF "train a diffusion model" config -> 113,673,219
    m = OracleUNet2D(sample_size=128, in_channels=3, out_channels=3, layers_per_block=2,
                     block_out_channels=(128, 128, 256, 256, 512, 512),
                     down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D",
                                       "AttnDownBlock2D", "DownBlock2D"),
                     up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D",
                                     "UpBlock2D"))
    assert _count(m) == 113_673_219


def test_param_count_c1_config():
    # BASELINE.json configs[0]: 64x64, 2 down / 2 up blocks -> 3,660,803
    m = OracleUNet2D(sample_size=64, block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2,
                     up_block_types=("UpBlock2D",) * 2)
    assert _count(m) == 3_660_803
    x = torch.randn(2, 3, 64, 64, generator=torch.manual_seed(1))
    with torch.no_grad():
        y = m(x, 999)[0]
    assert y.shape == x.shape and torch.isfinite(y).all()
