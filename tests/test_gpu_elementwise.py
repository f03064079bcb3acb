"""GPU parity: HBM-bound kernels vs the CPU oracle / torch fp32, through the C ABI."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def test_library_and_device():
    from drivescenegen_b200 import _lib
    lib = _lib.load()
    assert lib.dsg_version() >= 100
    assert lib.dsg_device_ok() == 1, "tests must run on an sm_100 device"


@pytest.mark.parametrize("n_steps", [1000, 750, 50])
def test_ddpm_step_bit_exact(n_steps):
    """dsg_ddpm_step == oracle step bit for bit on identical inputs (fp32 elementwise, no FMA contraction)."""
    from drivescenegen_b200 import ops
    from drivescenegen_b200.hostapi import DDPMScheduler
    from oracle.schedulers import OracleDDPMScheduler
    g = torch.Generator().manual_seed(3)
    shape = (2, 3, 32, 32)
    sch, osch = DDPMScheduler(), OracleDDPMScheduler()
    sch.set_timesteps(n_steps)
    osch.set_timesteps(n_steps)
    assert torch.equal(sch.timesteps.cpu(), osch.timesteps)
    tab = sch.coef_table(_dev())
    for t in [int(sch.timesteps[0]), int(sch.timesteps[len(sch.timesteps) // 2]), int(sch.timesteps[-2]), 0]:
        eps = torch.randn(shape, generator=g)
        x = torch.randn(shape, generator=g) * 1.5
        z = torch.randn(shape, generator=g)
        ref = osch.step(eps, t, x, variance_noise=z)
        got = ops.ddpm_step(eps.to(_dev()), x.to(_dev()), z.to(_dev()), tab, t)
        assert torch.equal(got.cpu(), ref), f"t={t}: max diff {(got.cpu() - ref).abs().max().item()}"
        # device-side row index (CUDA-graph path) gives the same bits
        row = torch.tensor([t], dtype=torch.int32, device=_dev())
        got2 = ops.ddpm_step(eps.to(_dev()), x.to(_dev()), z.to(_dev()), tab, 0, row_dev=row)
        assert torch.equal(got2, got)


def test_ddpm_scheduler_api_matches_oracle_with_generator():
    from drivescenegen_b200.hostapi import DDPMScheduler
    from oracle.schedulers import OracleDDPMScheduler
    sch, osch = DDPMScheduler(), OracleDDPMScheduler()
    sch.set_timesteps(750)
    osch.set_timesteps(750)
    g = torch.Generator().manual_seed(5)
    eps, x = torch.randn(1, 3, 16, 16, generator=g), torch.randn(1, 3, 16, 16, generator=g)
    # CPU generator + CUDA tensors: noise is drawn on the CPU and copied (upstream randn_tensor)
    got = sch.step(eps.to(_dev()), 500, x.to(_dev()), generator=torch.Generator().manual_seed(11)).prev_sample
    ref = osch.step(eps, 500, x, generator=torch.Generator().manual_seed(11))
    assert torch.equal(got.cpu(), ref)
    with pytest.raises(Exception):
        sch.step(eps, 500, x)  # CPU tensors: no CPU arithmetic path in the product


def test_ddim_step_bit_exact():
    from drivescenegen_b200 import ops
    from drivescenegen_b200.hostapi import DDIMScheduler
    from oracle.schedulers import OracleDDIMScheduler
    g = torch.Generator().manual_seed(4)
    sch, osch = DDIMScheduler(), OracleDDIMScheduler()
    sch.set_timesteps(50)
    osch.set_timesteps(50)
    assert torch.equal(sch.timesteps.cpu(), osch.timesteps)
    tab = sch.coef_table(_dev(), 0.0)
    for t in [980, 500, 20, 0]:
        eps, x = torch.randn(2, 3, 24, 24, generator=g), torch.randn(2, 3, 24, 24, generator=g)
        ref = osch.step(eps, t, x, eta=0.0)
        got = ops.ddpm_step(eps.to(_dev()), x.to(_dev()), None, tab, t, ddim=True)
        assert torch.equal(got.cpu(), ref), f"t={t}"


def test_ddpm_kat_on_gpu():
    """Upstream known-answer loop (258.9606 / 0.3372) driven through the CUDA kernel."""
    from drivescenegen_b200.hostapi import DDPMScheduler
    sch = DDPMScheduler()
    n = 4 * 3 * 8 * 8
    sample = (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2).contiguous().to(_dev())
    gen = torch.manual_seed(0)
    for t in reversed(range(len(sch))):
        residual = sample * t / (t + 1)
        sample = sch.step(residual, t, sample, generator=gen).prev_sample
    assert abs(sample.abs().sum().item() - 258.9606) < 1e-2
    assert abs(sample.abs().mean().item() - 0.3372) < 1e-3


def test_ddim_kat_on_gpu():
    from drivescenegen_b200.hostapi import DDIMScheduler
    sch = DDIMScheduler()
    sch.set_timesteps(10)
    n = 4 * 3 * 8 * 8
    sample = (torch.arange(n).reshape(3, 8, 8, 4) / n).permute(3, 0, 1, 2).contiguous().to(_dev())
    for t in sch.timesteps:
        residual = sample * t / (t + 1)
        sample = sch.step(residual, t, sample, eta=0.0).prev_sample
    assert abs(sample.abs().sum().item() - 172.0067) < 1e-2
    assert abs(sample.abs().mean().item() - 0.223967) < 1e-3


def test_add_noise_bit_exact():
    from drivescenegen_b200.hostapi import DDPMScheduler
    from oracle.schedulers import OracleDDPMScheduler
    g = torch.Generator().manual_seed(8)
    x0, nz = torch.rand(5, 3, 20, 20, generator=g) * 2 - 1, torch.randn(5, 3, 20, 20, generator=g)
    t = torch.randint(0, 1000, (5,), generator=g)
    ref = OracleDDPMScheduler().add_noise(x0, nz, t)
    got = DDPMScheduler().add_noise(x0.to(_dev()), nz.to(_dev()), t.to(_dev()))
    assert torch.equal(got.cpu(), ref)
    # ragged / empty
    e = DDPMScheduler().add_noise(x0[:0].to(_dev()), nz[:0].to(_dev()), t[:0].to(_dev()))
    assert e.shape[0] == 0


def test_add_noise_out_of_range_timestep_never_reads_out_of_bounds():
    """upstream: IndexError.  Host-visible timesteps raise; device-resident ones poison that sample with NaN."""
    from drivescenegen_b200.hostapi import DDPMScheduler
    s = DDPMScheduler()
    x0, nz = torch.ones(3, 3, 8, 8, device=_dev()), torch.ones(3, 3, 8, 8, device=_dev())
    with pytest.raises(IndexError):
        s.add_noise(x0, nz, torch.tensor([5, 1000, 7]))
    with pytest.raises(IndexError):
        s.add_noise(x0, nz, torch.tensor([-1, 0, 7]))
    got = s.add_noise(x0, nz, torch.tensor([5, 1000, 7], device=_dev()))
    assert torch.isfinite(got[0]).all() and torch.isnan(got[1]).all() and torch.isfinite(got[2]).all()


def test_scheduler_steps_propagate_nan_like_torch_clamp():
    """torch.clamp keeps NaN; a NaN model output (fp16 overflow inside the U-Net) must not be laundered into +-1."""
    from drivescenegen_b200 import ops
    from drivescenegen_b200.hostapi import DDIMScheduler, DDPMScheduler
    from oracle.schedulers import OracleDDIMScheduler, OracleDDPMScheduler
    g = torch.Generator().manual_seed(4)
    eps, x, z = (torch.randn(1, 3, 8, 8, generator=g) for _ in range(3))
    eps[0, 0, 0, 0] = float("nan")
    eps[0, 1, 2, 3] = float("inf")
    sch, osch = DDPMScheduler(), OracleDDPMScheduler()
    ref = osch.step(eps, 500, x, variance_noise=z)
    got = ops.ddpm_step(eps.to(_dev()), x.to(_dev()), z.to(_dev()), sch.coef_table(_dev()), 500).cpu()
    assert torch.isnan(ref[0, 0, 0, 0]) and torch.isnan(got[0, 0, 0, 0])
    assert torch.equal(torch.isnan(got), torch.isnan(ref)) and torch.equal(got.nan_to_num(0.0), ref.nan_to_num(0.0))
    dsch, dosch = DDIMScheduler(), OracleDDIMScheduler()
    dsch.set_timesteps(50)
    dosch.set_timesteps(50)
    refd = dosch.step(eps, 500, x)
    gotd = ops.ddpm_step(eps.to(_dev()), x.to(_dev()), None, dsch.coef_table(_dev()), 500, ddim=True).cpu()
    assert torch.equal(torch.isnan(gotd), torch.isnan(refd)) and torch.equal(gotd.nan_to_num(0.0), refd.nan_to_num(0.0))


def test_step_advance_walks_the_schedule_on_the_device():
    from drivescenegen_b200 import _lib
    from drivescenegen_b200._lib import check
    lib = _lib.load()
    sched = torch.tensor([980, 960, 3, 0], dtype=torch.int32, device=_dev())
    state = torch.tensor([0, 4], dtype=torch.int32, device=_dev())
    t_f = torch.zeros(5, device=_dev())
    row = torch.zeros(1, dtype=torch.int32, device=_dev())
    st = torch.cuda.current_stream().cuda_stream
    seen = []
    for _ in range(6):   # two calls past the end stay on the last entry
        check(lib.dsg_step_advance(sched.data_ptr(), state.data_ptr(), t_f.data_ptr(), 5, row.data_ptr(), st), "advance")
        seen.append((int(row.item()), t_f.tolist(), int(state[0].item())))
    assert [s[0] for s in seen] == [980, 960, 3, 0, 0, 0]
    assert all(s[1] == [float(s[0])] * 5 for s in seen)
    assert [s[2] for s in seen] == [1, 2, 3, 4, 5, 6]


def test_latent_to_image_exact():
    from drivescenegen_b200 import ops
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(2, 3, 16, 24, generator=g) * 1.2
    ref = (lat / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
    ref_u8 = (ref * 255).round().astype("uint8")
    u8, f32 = ops.latent_to_image(lat.to(_dev()))
    assert (f32.cpu().numpy() == ref).all()
    assert (u8.cpu().numpy() == ref_u8).all()


def test_time_embed_vs_oracle():
    from drivescenegen_b200 import ops
    from oracle.unet import TimestepEmbedding, timestep_embedding
    torch.manual_seed(0)
    te = TimestepEmbedding(64, 256)
    projs = [torch.nn.Linear(256, c) for c in (64, 128, 512)]
    t = torch.tensor([999, 0, 17, 500], dtype=torch.long)
    with torch.no_grad():
        emb = te(timestep_embedding(t, 64))
        ref = torch.cat([p(F.silu(emb)) for p in projs], dim=1)
    half = 32
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    d = _dev()
    wp = torch.cat([p.weight for p in projs], 0).detach()
    bp = torch.cat([p.bias for p in projs], 0).detach()
    out, _ = ops.time_embed(t.float().to(d), freqs.to(d), True, te.linear_1.weight.detach().to(d),
                            te.linear_1.bias.detach().to(d), te.linear_2.weight.detach().to(d),
                            te.linear_2.bias.detach().to(d), wp.to(d), bp.to(d))
    # fp32 both sides; tolerance covers summation order and sin/cos argument reduction at t ~ 1000
    assert torch.allclose(out.cpu(), ref, atol=2e-4, rtol=1e-4), (out.cpu() - ref).abs().max()


def test_conv_in_out_vs_torch():
    from drivescenegen_b200 import ops
    torch.manual_seed(1)
    d = _dev()
    x = torch.randn(2, 3, 20, 28)
    ci = torch.nn.Conv2d(3, 64, 3, padding=1)
    with torch.no_grad():
        ref = ci(x)
    got = ops.conv_in(x.to(d), ci.weight.detach().to(d), ci.bias.detach().to(d))
    got = got.float().permute(0, 3, 1, 2).cpu()
    assert torch.allclose(got, ref, atol=3e-3, rtol=2e-3), (got - ref).abs().max()  # fp16 output rounding
    co = torch.nn.Conv2d(64, 3, 3, padding=1)
    a = torch.randn(2, 20, 28, 64).half()
    with torch.no_grad():
        ref = co(a.float().permute(0, 3, 1, 2))
    got = ops.conv_out(a.to(d), co.weight.detach().to(d), co.bias.detach().to(d)).cpu()
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()  # fp32 math on both sides


@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 3, 64, 20, 28), (3, 3, 64, 64, 64), (1, 3, 64, 13, 37), (2, 1, 64, 8, 8),
                                            (16, 3, 64, 256, 256), (2, 3, 64, 9, 512), (2, 3, 128, 16, 16),
                                            (2, 4, 64, 16, 16)])
def test_conv_in_stats_tensor_core_form_vs_torch(n, cin, cout, h, w):
    """dsg_conv_in_stats: mma.sync form (cout 64, cin*9 <= 32; fp16 operands, fp32 accumulate) and the fall-back for
    other widths, against torch fp32 conv2d; the fused statistics against sums of the fp32 reference output."""
    from drivescenegen_b200 import ops
    torch.manual_seed(3)
    d = _dev()
    x = torch.randn(n, cin, h, w)
    ci = torch.nn.Conv2d(cin, cout, 3, padding=1)
    with torch.no_grad():
        ref = ci.to(d)(x.to(d)).float().cpu()
    got, st = ops.conv_in_stats(x.to(d), ci.weight.detach().to(d), ci.bias.detach().to(d))
    got = got.float().permute(0, 3, 1, 2).cpu()
    # fp16 rounding of x, w (2^-11 each, 27 terms) and of the output
    assert torch.allclose(got, ref, atol=6e-3, rtol=4e-3), (got - ref).abs().max()
    assert ((got - ref).norm() / ref.norm()).item() < 1.5e-3
    st = st.cpu().double()
    s_ref = ref.double().sum(dim=(2, 3))
    q_ref = (ref.double() ** 2).sum(dim=(2, 3))
    s_got, q_got = st[..., 0] / 2 ** 24, st[..., 1] / 2 ** 20
    npx = h * w
    assert (s_got - s_ref).abs().max().item() <= 2e-3 * npx ** 0.5 + 1e-3 * s_ref.abs().max().item()
    assert ((q_got - q_ref).abs() / q_ref).max().item() < 5e-3
    # and they agree with a separate statistics pass over the stored fp16 tensor
    got16, _ = ops.conv_in_stats(x.to(d), ci.weight.detach().to(d), ci.bias.detach().to(d))
    st2 = ops.gn_stats(got16).cpu().double()
    assert ((st2[..., 1] - st[..., 1]).abs() / st[..., 1]).max().item() < 2e-3


@pytest.mark.parametrize("n,h,w,cout", [(2, 20, 28, 3), (1, 64, 64, 3), (2, 9, 70, 3), (3, 17, 130, 4), (1, 8, 16, 1),
                                        (4, 256, 256, 3)])
def test_conv_out_fused_groupnorm_silu_vs_torch(n, h, w, cout):
    """dsg_conv_out_fused (conv_norm_out + SiLU + conv_out in one mma.sync pass over the raw tensor) against torch fp32
    group_norm -> silu -> conv2d, and its already-activated form (coef = NULL) against conv2d alone; also against the
    unfused library path (gn_apply + conv_out), which stages bit-identical fp16 activations."""
    import torch.nn.functional as F
    from drivescenegen_b200 import ops
    torch.manual_seed(4)
    d = _dev()
    x = (torch.randn(n, h, w, 64) * 1.5 + 0.3).half()
    gamma, beta = 1 + 0.2 * torch.randn(64), 0.2 * torch.randn(64)
    co = torch.nn.Conv2d(64, cout, 3, padding=1)
    wt, bs = co.weight.detach(), co.bias.detach()
    xn = x.float().permute(0, 3, 1, 2)
    with torch.no_grad():
        act_ref = F.silu(F.group_norm(xn, 32, gamma, beta, 1e-5))
        ref = F.conv2d(act_ref, wt, bs, padding=1)
        ref_plain = F.conv2d(xn, wt, bs, padding=1)
    xd = x.to(d)
    st = ops.gn_stats(xd)
    coef = ops.gn_coef(st, None, gamma.to(d), beta.to(d), 32, 1e-5, h * w)
    got = ops.conv_out_fused(xd, coef, wt.to(d), bs.to(d)).cpu()
    assert got.shape == ref.shape
    assert ((got - ref).norm() / ref.norm()).item() < 2e-3
    assert torch.allclose(got, ref, atol=8e-3, rtol=5e-3), (got - ref).abs().max()
    got_plain = ops.conv_out_fused(xd, None, wt.to(d), bs.to(d)).cpu()
    assert ((got_plain - ref_plain).norm() / ref_plain.norm()).item() < 1e-3
    # the unfused path: same fp16 activations, fp32 weights on CUDA cores -> differs only by weight rounding / sum order
    act = ops.group_norm(xd, None, gamma.to(d), beta.to(d), 32, 1e-5, 1, stats1=st)
    unf = ops.conv_out(act, wt.to(d), bs.to(d)).cpu()
    assert ((got - unf).norm() / unf.norm()).item() < 1e-3


@pytest.mark.parametrize("c1,c2,hw", [(64, 0, (32, 32)), (128, 64, (16, 16)), (512, 256, (8, 8)), (256, 128, (12, 20)),
                                      (1024, 0, (8, 8))])
@pytest.mark.parametrize("act", [0, 1])
def test_group_norm_vs_torch(c1, c2, hw, act):
    from drivescenegen_b200 import ops
    torch.manual_seed(2)
    d = _dev()
    n = 3
    x1 = (torch.randn(n, *hw, c1) * 2 + 0.7).half()
    x2 = (torch.randn(n, *hw, c2) * 0.5 - 3.0).half() if c2 else None  # large mean offset: cancellation check
    gamma, beta = torch.randn(c1 + c2), torch.randn(c1 + c2)
    cat = x1 if x2 is None else torch.cat([x1, x2], dim=3)
    ref = F.group_norm(cat.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    if act:
        ref = F.silu(ref)
    got = ops.group_norm(x1.to(d), None if x2 is None else x2.to(d), gamma.to(d), beta.to(d), 32, 1e-5, act)
    got = got.float().permute(0, 3, 1, 2).cpu()
    assert torch.allclose(got, ref, atol=6e-3, rtol=3e-3), (got - ref).abs().max()  # fp16 output rounding


@pytest.mark.parametrize("tokens,heads,hd", [(256, 8, 8), (1024, 64, 8), (200, 4, 16), (640, 2, 64)])
def test_attention_vs_torch(tokens, heads, hd):
    from drivescenegen_b200 import ops
    torch.manual_seed(3)
    d = _dev()
    n, c = 2, heads * hd
    qkv = (torch.randn(n, tokens, 3 * c) * 1.5).half()
    q, k, v = [t.float().view(n, tokens, heads, hd).transpose(1, 2) for t in qkv.split(c, dim=2)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, tokens, c)
    got = ops.attention(qkv.to(d), heads, hd).float().cpu()
    assert torch.allclose(got, ref, atol=3e-3, rtol=3e-3), (got - ref).abs().max()
