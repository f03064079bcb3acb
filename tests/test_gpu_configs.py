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


def test_run_from_host_pipelined_equals_serial_loop():
    """DenoiseSession.run_from_host (copies on side streams) gives bit-identical samples to the serial
    copy-in / step / copy-out loop, and hands every step's result to the host exactly once, in order."""
    from drivescenegen_b200.hostapi import DDPMScheduler, DenoiseSession, UNet2DModel
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = UNet2DModel(sample_size=64, block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2,
                        up_block_types=("UpBlock2D",) * 2).to(dev).eval()
    sched = DDPMScheduler()
    sched.set_timesteps(1000)
    shape = (2, 3, 64, 64)
    sess = DenoiseSession(model, sched, shape)
    gen = torch.Generator().manual_seed(5)
    x0 = torch.randn(shape, generator=gen)
    noise = [torch.randn(shape, generator=gen).pin_memory() for _ in range(3)]
    ts = [int(t) for t in sched.timesteps[:7]]

    out = torch.empty(shape).pin_memory()
    serial = []
    sess.x.copy_(x0)
    for i, t in enumerate(ts):
        serial.append(sess.step_from_host(t, noise[i % 3], out).clone())

    outs = [torch.empty(shape).pin_memory() for _ in range(2)]
    got = []
    for n in (len(ts), 1, 0):      # full run, a single step, an empty run
        got.clear()
        sess.x.copy_(x0)
        sess.run_from_host(ts[:n], noise, outs, on_result=lambda i, o: got.append((i, o.clone())))
        assert [i for i, _ in got] == list(range(n))
        for (i, o), ref in zip(got, serial):
            assert torch.equal(o, ref), f"step {i} differs"
    with pytest.raises(ValueError):
        sess.run_from_host(ts, noise, outs[:1])


@pytest.mark.parametrize("group", [1, 2, 4])
def test_l2_sample_groups_are_bit_identical_to_whole_batch_launches(group):
    """engine.l2_group re-issues the full-resolution ops sample-group by sample-group (depth first) so consumers find
    their inputs in L2; every op of the U-Net is independent per sample, so the output must not change by a bit."""
    from drivescenegen_b200.hostapi import UNet2DModel
    torch.manual_seed(0)
    model = UNet2DModel(sample_size=64, **REF_CFG).to(_dev()).eval()
    x = torch.randn(8, 3, 64, 64, generator=torch.Generator().manual_seed(9)).to(_dev())
    eng = model.engine()
    eng.l2_group = 0
    eng.programs.clear()
    with torch.no_grad():
        ref = model(x, 321).sample.clone()
        n0 = eng.program(8, 64, 64).n_launches
        eng.l2_group = group
        eng.programs.clear()
        got = model(x, 321).sample.clone()
        n1 = eng.program(8, 64, 64).n_launches
    eng.l2_group = 0
    eng.programs.clear()
    assert n1 > n0
    assert torch.equal(got, ref)
