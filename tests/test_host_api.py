"""CPU tests: host logic of the reference-facing API, C-ABI symbol table, shims, plumbing of the reference scripts."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built_lib():
    from drivescenegen_b200.build import build_library
    return build_library()


def test_cabi_exports_every_declared_symbol(built_lib):
    """The library loads without a GPU and exports exactly what include/dsg_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "dsg_b200.h")).read()
    declared = set(re.findall(r"\b(dsg_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("dsg_conv_args")
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in dsg_b200.h but not exported"
    from drivescenegen_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    l = _lib.load()
    assert l.dsg_version() == 100
    assert l.dsg_packed_k(0, 64, 128) == 9 * 64 + 128 and l.dsg_packed_k(2, 64, 0) == 256
    assert l.dsg_packed_rows(2, 64) == 256
    # host-side planning of the batched weight re-pack: one block per output channel for the forward layouts, one per
    # 128 co x 8 ci (64 ci for 1x1) tile for the data-gradient layouts, -1 when a source row does not fit the staging
    assert l.dsg_pack_job_blocks(0, 512, 1024, 1024) == 512
    assert l.dsg_pack_job_blocks(2, 128, 128, 0) == 128
    assert l.dsg_pack_job_blocks(10, 200, 72, 0) == 9 * 2 and l.dsg_pack_job_blocks(13, 512, 512, 0) == 8 * 4
    assert l.dsg_pack_job_blocks(0, 64, 2048, 0) == -1 and l.dsg_pack_job_blocks(5, 64, 64, 0) == -1
    # replays of a captured graph are reported by the host: the counter is the number of kernels executed
    n0 = l.dsg_launch_count()
    l.dsg_count_graph_launches(475)
    l.dsg_count_graph_launches(-3)
    assert l.dsg_launch_count() - n0 == 475


def test_conv_args_struct_layout_matches_header(tmp_path, built_lib):
    """sizeof/offsetof of dsg_conv_args as compiled by gcc == the ctypes mirror."""
    src = tmp_path / "lay.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dsg_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(dsg_conv_args),offsetof(dsg_conv_args,x),offsetof(dsg_conv_args,wpacked),'
                   'offsetof(dsg_conv_args,residual),offsetof(dsg_conv_args,impl));return 0;}\n')
    exe = tmp_path / "lay"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    from drivescenegen_b200._lib import ConvArgs
    exp = [ctypes.sizeof(ConvArgs), ConvArgs.x.offset, ConvArgs.wpacked.offset, ConvArgs.residual.offset,
           ConvArgs.impl.offset]
    assert got == exp


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "drivescenegen_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports oracle"
    for f in ("shims/diffusers/__init__.py", "shims/accelerate/__init__.py"):
        assert "oracle" not in open(os.path.join(ROOT, f)).read()


def test_cpu_tensors_fail_loudly():
    from drivescenegen_b200.hostapi import DDPMScheduler, UNet2DModel
    from drivescenegen_b200._lib import DsgError
    m = UNet2DModel(sample_size=64, block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2,
                    up_block_types=("UpBlock2D",) * 2)
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(DsgError):
        m(x, 0)
    with pytest.raises(DsgError):
        DDPMScheduler().step(x, 10, x)
    with pytest.raises(DsgError):
        DDPMScheduler().add_noise(x, x, torch.tensor([1]))


def test_unet_ctor_validation_and_state_dict_keys():
    from drivescenegen_b200.hostapi import UNet2DModel
    from oracle.unet import OracleUNet2D
    with pytest.raises(ValueError):
        UNet2DModel(block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2, up_block_types=("UpBlock2D",))
    with pytest.raises(ValueError):
        UNet2DModel(block_out_channels=(64, 128), down_block_types=("NopeBlock",) * 2,
                    up_block_types=("UpBlock2D",) * 2)
    cfg = dict(block_out_channels=(128, 128, 256, 256, 512, 512), layers_per_block=2,
               down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
               up_block_types=("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4)
    m, o = UNet2DModel(sample_size=128, **cfg), OracleUNet2D(sample_size=128, **cfg)
    assert sum(p.numel() for p in m.parameters()) == 113_673_219
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert m.config.sample_size == 128 and m.config["in_channels"] == 3
    assert m.dtype == torch.float32 and m.device.type == "cpu"


def test_scheduler_host_logic_matches_oracle():
    from drivescenegen_b200.hostapi import DDIMScheduler, DDPMScheduler
    from oracle.schedulers import OracleDDPMScheduler
    s, o = DDPMScheduler(), OracleDDPMScheduler()
    assert s.num_train_timesteps == 1000 and len(s) == 1000 and s.config.variance_type == "fixed_small"
    for n in (1000, 750, 50, 7):
        s.set_timesteps(n)
        o.set_timesteps(n)
        assert torch.equal(s.timesteps, o.timesteps)
    with pytest.raises(ValueError):
        s.set_timesteps(1001)
    s.set_timesteps(750)
    o.set_timesteps(750)
    # coefficient row == what the oracle's step uses (checked through a 1-element step)
    for t in (749, 400, 1, 0):
        r = s._coef_row(t)
        x, e, z = torch.tensor([0.3]), torch.tensor([-0.7]), torch.tensor([1.1])
        ref = o.step(e, t, x, variance_noise=z)
        c = [torch.tensor(v, dtype=torch.float32) for v in r]
        x0 = ((x - c[0] * e) / c[1]).clamp(-c[5], c[5])
        mine = c[2] * x0 + c[3] * x
        if r[6]:
            mine = mine + c[4] * z
        assert torch.equal(mine, ref)
    d = DDIMScheduler()
    d.set_timesteps(50)
    assert d.timesteps[0].item() == 980 and d.timesteps[-1].item() == 0
    # DDPM -> DDIM swap via from_config ignores keys DDIM does not take
    d2 = DDIMScheduler.from_config(s.config)
    assert d2.config.num_train_timesteps == 1000


def test_scheduler_and_pipeline_config_round_trip(tmp_path):
    from drivescenegen_b200.hostapi import DDPMPipeline, DDPMScheduler, UNet2DModel
    m = UNet2DModel(sample_size=(64, 64), block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2,
                    up_block_types=("UpBlock2D",) * 2)
    p = DDPMPipeline(unet=m, scheduler=DDPMScheduler())
    p.save_pretrained(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["model_index.json", "scheduler", "unet"]
    assert os.path.isfile(tmp_path / "unet" / "diffusion_pytorch_model.bin")
    assert os.path.isfile(tmp_path / "unet" / "config.json")
    assert os.path.isfile(tmp_path / "scheduler" / "scheduler_config.json")
    p2 = DDPMPipeline.from_pretrained(str(tmp_path), variant="fp16")
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), p2.unet.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert tuple(p2.unet.config.sample_size) == (64, 64)
    m.save_pretrained(str(tmp_path / "st"), safe_serialization=True, variant="fp16")
    m3 = UNet2DModel.from_pretrained(str(tmp_path / "st"), variant="fp16")
    assert torch.equal(m3.conv_in.weight, m.conv_in.weight)


def test_cosine_schedule():
    from drivescenegen_b200.hostapi import get_cosine_schedule_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1.0)
    sch = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=4, num_training_steps=12)
    lrs = []
    for _ in range(12):
        lrs.append(sch.get_last_lr()[0])
        opt.step()
        sch.step()
    assert lrs[0] == 0.0 and abs(lrs[2] - 0.5) < 1e-12 and abs(lrs[4] - 1.0) < 1e-12
    assert abs(lrs[8] - 0.5) < 1e-9 and lrs[11] < 0.05


def test_shims_export_reference_import_surface():
    code = ("import diffusers, accelerate\n"
            "from diffusers import UNet2DModel, DDPMScheduler, DDPMPipeline\n"
            "from diffusers.optimization import get_cosine_schedule_with_warmup\n"
            "from accelerate import Accelerator, notebook_launcher\n"
            "print(diffusers.__version__, accelerate.__version__)\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "shims"))
    out = subprocess.check_output([sys.executable, "-c", code], env=env, cwd="/tmp").decode()
    assert out.split() == ["0.20.0", "0.22.0"]
