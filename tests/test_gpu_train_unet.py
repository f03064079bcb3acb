"""GPU parity of the whole training path: UNet2DModel forward + hand-written backward vs the CPU oracle under torch
autograd (fp32).  Tolerance (fp16 operands / activations / activation gradients, fp32 accumulation): relative L2 error of
each parameter gradient <= 1e-2, of all gradients together <= 5e-3, of the output <= 3e-3 (measured on B200: worst single
gradient ~4e-3, all together ~2e-3, output ~1e-3), up to the benchmarked 256x256 resolution."""
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG_C1 = dict(sample_size=64, block_out_channels=(64, 128), down_block_types=("DownBlock2D",) * 2,
              up_block_types=("UpBlock2D",) * 2)
CFG_REF = dict(sample_size=64, block_out_channels=(64, 128, 256, 512), down_block_types=("DownBlock2D",) * 4,
               up_block_types=("UpBlock2D",) * 4)


def _dev():
    return torch.device("cuda", 0)


def _pair(cfg, seed=0):
    from drivescenegen_b200.hostapi import UNet2DModel
    from oracle.unet import OracleUNet2D
    torch.manual_seed(seed)
    oracle = OracleUNet2D(**cfg)
    # make every parameter matter: perturb GroupNorm affines / biases away from their 1 / 0 initial values
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    model = UNet2DModel(**cfg)
    model.load_state_dict(oracle.state_dict())
    return oracle, model.to(_dev())


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name,cfg,size,batch", [("c1", CFG_C1, 64, 2), ("ref", CFG_REF, 64, 2), ("ref128", CFG_REF, 128, 1),
                                                 ("ref256", CFG_REF, 256, 1)])
def test_unet_backward_matches_oracle_autograd(name, cfg, size, batch):
    oracle, model = _pair(cfg)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(batch, 3, size, size, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    target = torch.randn(batch, 3, size, size, generator=g)
    ref_out = oracle(x, t)[0]
    ref_loss = torch.nn.functional.mse_loss(ref_out, target)
    ref_loss.backward()
    out = model(x.to(_dev()), t.to(_dev()), return_dict=False)[0]
    assert out.requires_grad
    loss = torch.nn.functional.mse_loss(out, target.to(_dev()))
    loss.backward()
    rel_out = _rel(out.detach().cpu(), ref_out.detach())
    assert rel_out < 3e-3, rel_out
    ref = dict(oracle.named_parameters())
    num = den = 0.0
    worst = []
    total = sum(p.grad.pow(2).sum().item() for p in ref.values()) ** 0.5
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        assert torch.isfinite(p.grad).all(), n
        gr, gg = ref[n].grad, p.grad.cpu()
        num += (gg - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
        if gr.norm().item() < 1e-5 * total:
            # e.g. attention to_k.bias: softmax is invariant to a constant key shift, the true gradient is zero
            assert (gg - gr).norm().item() < 1e-4 * total, n
            continue
        worst.append((_rel(gg, gr), n))
    worst.sort(reverse=True)
    print(f"[parity] backward {name} {size}x{size} b={batch}: out rel_l2={rel_out:.2e} all-grads rel_l2="
          f"{(num / den) ** 0.5:.2e} worst grad {worst[0][1]} {worst[0][0]:.2e}", file=sys.stderr)
    assert (num / den) ** 0.5 < 5e-3, ((num / den) ** 0.5, worst[:5])
    assert worst[0][0] < 1e-2, worst[:8]


def test_unet_backward_no_loss_scale_and_large_loss_scale_agree():
    """the internal power-of-two gradient scale makes the result independent of the caller's loss scale."""
    _, model = _pair(CFG_C1)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 64, 64, generator=g).to(_dev())
    target = torch.randn(2, 3, 64, 64, generator=g).to(_dev())
    grads = []
    for scale in (1.0, 65536.0):
        model.zero_grad(set_to_none=True)
        out = model(x, 500, return_dict=False)[0]
        (torch.nn.functional.mse_loss(out, target) * scale).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]) / scale)
    assert torch.equal(grads[0], grads[1])   # power-of-two scales: bit-identical


def test_unet_backward_is_deterministic_and_accumulates():
    _, model = _pair(CFG_C1)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, 64, 64, generator=g).to(_dev())
    target = torch.randn(2, 3, 64, 64, generator=g).to(_dev())

    def grads():
        out = model(x, torch.tensor([10, 900], device=_dev()), return_dict=False)[0]
        torch.nn.functional.mse_loss(out, target).backward()
        return torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    model.zero_grad(set_to_none=True)
    a = grads()
    model.zero_grad(set_to_none=True)
    b = grads()
    assert torch.equal(a, b)
    c = grads()   # no zero_grad: gradients accumulate like any autograd graph
    assert torch.allclose(c, 2 * a, rtol=1e-6, atol=0)


def test_training_step_reduces_loss_with_torch_adamw():
    """a few optimizer steps through the public API (forward, backward, torch.optim.AdamW) on a fixed batch."""
    _, model = _pair(CFG_C1)
    model.train()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4, 3, 64, 64, generator=g).to(_dev())
    target = torch.randn(4, 3, 64, 64, generator=g).to(_dev())
    t = torch.randint(0, 1000, (4,), generator=g).to(_dev())
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    losses = []
    for _ in range(6):
        out = model(x, t, return_dict=False)[0]
        loss = torch.nn.functional.mse_loss(out, target)
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


def _train_loop(model, accelerator, optimizer, lr_sched, sched, batches, fused_expected):
    """the call sequence of DriveSceneGen/pipeline/training_pipeline.py:72-91 (train_loop body)."""
    import torch.nn.functional as F
    model, optimizer, lr_sched = accelerator.prepare(model, optimizer, lr_sched)
    losses = []
    for clean_images in batches:
        noise = torch.randn(clean_images.shape, generator=torch.Generator().manual_seed(len(losses))).to(clean_images.device)
        bs = clean_images.shape[0]
        timesteps = torch.randint(0, sched.num_train_timesteps, (bs,), generator=torch.Generator().manual_seed(99),
                                  ).long().to(clean_images.device)
        noisy_images = sched.add_noise(clean_images, noise, timesteps)
        with accelerator.accumulate(model):
            noise_pred = model(noisy_images, timesteps, return_dict=False)[0]
            loss = F.mse_loss(noise_pred, noise)
            accelerator.backward(loss)
            accelerator.clip_grad_norm_(model.parameters(), 1.0)
            optimizer.step()
            lr_sched.step()
            optimizer.zero_grad()
        losses.append(loss.detach().item())
    assert bool(accelerator._fused_state) == fused_expected
    return losses


def test_training_graphs_are_bit_identical_to_launch_by_launch(monkeypatch):
    """From its second step on a training program replays its forward and backward as CUDA graphs (same kernels, same
    order, static buffers): parameters and losses after 6 steps are bit-identical to the launch-by-launch path, the graphs
    really are in use, and their replays are reported to the library's launch counter."""
    from drivescenegen_b200 import _lib
    from drivescenegen_b200.hostapi import Accelerator, DDPMScheduler, get_cosine_schedule_with_warmup
    g = torch.Generator().manual_seed(22)
    batches = [torch.rand(4, 3, 64, 64, generator=g).mul(2).sub(1).to(_dev()) for _ in range(6)]
    results = []
    for graphs in ("1", "0"):
        monkeypatch.setenv("DSG_TRAIN_GRAPH", graphs)   # read when a TrainProgram is built
        _, model = _pair(CFG_REF, seed=4)
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        lr_sched = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=2, num_training_steps=8)
        acc = Accelerator(mixed_precision="fp16", gradient_accumulation_steps=1)
        n0 = _lib.launch_count()
        losses = _train_loop(model, acc, opt, lr_sched, DDPMScheduler(), batches, True)
        launches = _lib.launch_count() - n0
        progs = list(model.engine(train=True).train_programs.values())
        captured = [p for p in progs if p._fwd_graph is not None]
        assert bool(captured) == (graphs == "1")
        results.append((losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone(), launches))
    (l_g, p_g, n_g), (l_e, p_e, n_e) = results
    assert l_g == l_e, (l_g, l_e)
    assert torch.equal(p_g, p_e)
    # the graph run executes the same kernels + one extra pass through the capture (which launches nothing)
    assert n_g >= n_e, (n_g, n_e)


@pytest.mark.parametrize("precision", ["fp16", "no"])
def test_reference_train_loop_sequence_fused_vs_torch_optimizer(precision):
    """The reference's training-loop call sequence through the shims: the fused flat-buffer clip + AdamW path (what
    `accelerator.prepare` sets up on CUDA) gives the same parameters as torch.optim.AdamW + clip_grad_norm_ on the
    same gradients."""
    from drivescenegen_b200.hostapi import Accelerator, DDPMScheduler, get_cosine_schedule_with_warmup
    g = torch.Generator().manual_seed(21)
    batches = [torch.rand(4, 3, 64, 64, generator=g).mul(2).sub(1).to(_dev()) for _ in range(4)]
    results = []
    for fused in (True, False):
        _, model = _pair(CFG_C1, seed=3)
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        lr_sched = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=2, num_training_steps=8)
        acc = Accelerator(mixed_precision=precision, gradient_accumulation_steps=1)
        if not fused:
            acc._flatten_parameters = lambda m: None   # keeps the generic torch path (GradScaler.step + AdamW)
        losses = _train_loop(model, acc, opt, lr_sched, DDPMScheduler(), batches, fused)
        results.append((losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()))
    (l_f, p_f), (l_t, p_t) = results
    assert all(abs(a - b) <= 2e-3 * abs(b) for a, b in zip(l_f, l_t)), (l_f, l_t)
    # AdamW normalises the update (|step| ~ lr whatever the gradient's size): parameters move by <= ~3.5e-4 in these 4
    # steps.  The two paths round the unscale/clip differently, so parameters whose true gradient is zero (pure
    # rounding noise, e.g. attention to_k.bias) may step in different directions; everything else agrees to ~1e-6.
    diff = (p_f - p_t).abs()
    assert diff.max().item() < 4e-4, diff.max().item()
    assert (diff > 4e-6).float().mean().item() < 2e-3, (diff > 4e-6).float().mean().item()
    assert l_f[-1] < l_f[0]


def test_full_size_training_step_properties_batch32_256():
    """BASELINE configs[2] at its full size (reference U-Net, 256x256, batch 32), through properties that need no CPU
    oracle: (1) two identical forward+backward passes give bit-identical gradients (every kernel is deterministic);
    (2) the gradient of the mean loss over 32 samples is the mean of the gradients of its two 16-sample halves
    (the backward is linear in the per-sample losses; GroupNorm / attention never mix samples)."""
    import torch.nn.functional as F
    from drivescenegen_b200.hostapi import UNet2DModel
    cfg = dict(CFG_REF, sample_size=256)
    torch.manual_seed(0)
    model = UNet2DModel(**cfg).to(_dev()).train()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(32, 3, 256, 256, generator=g).to(_dev())
    noise = torch.randn(32, 3, 256, 256, generator=g).to(_dev())
    t = torch.randint(0, 1000, (32,), generator=g).to(_dev())

    def grads(sl):
        model.zero_grad(set_to_none=True)
        out = model(x[sl], t[sl], return_dict=False)[0]
        loss = F.mse_loss(out, noise[sl])
        loss.backward()
        return torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone(), loss.item()

    full1, l1 = grads(slice(0, 32))
    full2, l2 = grads(slice(0, 32))
    assert torch.isfinite(full1).all() and full1.abs().max().item() > 0
    assert l1 == l2 and torch.equal(full1, full2)
    ga, la = grads(slice(0, 16))
    gb, lb = grads(slice(16, 32))
    assert abs(0.5 * (la + lb) - l1) <= 1e-5 * abs(l1)
    # each half runs with its own power-of-two gradient scale and fp16 roundings: agreement at fp16 resolution
    assert _rel(0.5 * (ga + gb), full1) < 5e-3


def test_early_gradient_slice_is_complete_at_the_backward_boundary():
    """Data-parallel overlap: the flat gradient buffer is laid out [late ..., early ...] and the backward
    program fires a hook once it has crossed from the mid block into the down path.  At that moment every gradient of the
    early slice (up blocks, mid block, conv_norm_out, conv_out, minus the time_emb_proj layers) must already hold its
    FINAL value — that is what lets the all-reduce of that slice run under the rest of the backward pass."""
    import torch.nn.functional as F
    from drivescenegen_b200.hostapi.training import flat_layout
    _, model = _pair(CFG_REF)
    model.train()
    order, early_off = flat_layout(model)
    names = [n for n, _ in order]
    first_early = sum(1 for n, p in order if sum(q.numel() for _, q in order[:names.index(n)]) < early_off)
    assert all(n.startswith(("up_blocks.", "mid_block.", "conv_norm_out.", "conv_out.")) and "time_emb_proj" not in n
               for n in names[first_early:])
    assert all("time_emb_proj" in n or n.startswith(("conv_in.", "time_embedding.", "down_blocks."))
               for n in names[:first_early])
    total = sum(p.numel() for _, p in order)
    assert 0.5 < (total - early_off) / total < 0.8          # most of the buffer can hide under the down path
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 3, 64, 64, generator=g).to(_dev())
    target = torch.randn(2, 3, 64, 64, generator=g).to(_dev())
    t = torch.tensor([7, 911], device=_dev())

    def run():
        model.zero_grad(set_to_none=True)
        F.mse_loss(model(x, t, return_dict=False)[0], target).backward()
    run()                                                  # builds the program
    fg = model._flat_grads
    seen = {}

    def hook(prog):
        seen["early"] = fg.flat[prog.slot][fg.early_offset:].clone()      # stream-ordered snapshot at the boundary
        seen["late"] = fg.flat[prog.slot][:fg.early_offset].clone()
    for prog in model._engine.train_programs.values():
        prog.on_early_ready = hook
    fg.flat[0].fill_(float("nan"))                         # nothing may survive from the first pass
    fg.flat[1].fill_(float("nan"))
    run()
    torch.cuda.synchronize()
    final = fg.flat[fg.last]
    assert "early" in seen, "the backward program never reached its boundary op"
    assert torch.isfinite(final).all()
    assert torch.equal(seen["early"], final[fg.early_offset:])
    assert not torch.equal(seen["late"].nan_to_num(0.0), final[:fg.early_offset])   # the late slice was still being written
    # .grad tensors are views of the flat buffer in layout order
    off = 0
    for n, p in order:
        assert p.grad.data_ptr() == final.data_ptr() + 4 * off, n
        off += p.numel()


def test_fused_adamw_state_is_visible_in_state_dict_and_resumes_bit_exactly():
    """the fused AdamW path keeps its moments in flat buffers; optimizer.state / state_dict() must expose them
    (checkpoint) and load_state_dict() must restore them (resume): an interrupted-and-resumed run equals an
    uninterrupted one bit for bit."""
    from drivescenegen_b200.hostapi import Accelerator, DDPMScheduler, get_cosine_schedule_with_warmup
    g = torch.Generator().manual_seed(22)
    batches = [torch.rand(2, 3, 64, 64, generator=g).mul(2).sub(1).to(_dev()) for _ in range(4)]

    def fresh():
        _, model = _pair(CFG_C1, seed=4)
        model.train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        lrs = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=1, num_training_steps=8)
        return model, opt, lrs, Accelerator(mixed_precision="no")

    model, opt, lrs, acc = fresh()
    _train_loop(model, acc, opt, lrs, DDPMScheduler(), batches, True)
    want = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()

    model, opt, lrs, acc = fresh()
    _train_loop(model, acc, opt, lrs, DDPMScheduler(), batches[:2], True)
    wrapped = acc._optimizers[0]
    sd = wrapped.state_dict()
    assert len(sd["state"]) == len(list(model.parameters()))
    st0 = sd["state"][0]
    assert float(st0["step"]) == 2.0 and st0["exp_avg"].abs().sum().item() > 0 and st0["exp_avg_sq"].min().item() >= 0
    p0 = next(iter(model.parameters()))
    assert wrapped.state[p0]["exp_avg"].shape == p0.shape
    ckpt = {"model": {k: v.clone() for k, v in model.state_dict().items()},
            "opt": {"state": {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in sd["state"].items()},
                    "param_groups": sd["param_groups"]},
            "lrs": lrs.state_dict()}

    model2, opt2, lrs2, acc2 = fresh()
    model2.load_state_dict(ckpt["model"])
    model2, opt2w, lrs2w = acc2.prepare(model2, opt2, lrs2)
    opt2w.load_state_dict(ckpt["opt"])
    lrs2w.load_state_dict(ckpt["lrs"])
    sched = DDPMScheduler()
    import torch.nn.functional as F
    for i, clean in enumerate(batches[2:], start=2):
        noise = torch.randn(clean.shape, generator=torch.Generator().manual_seed(i)).to(clean.device)
        ts = torch.randint(0, 1000, (clean.shape[0],), generator=torch.Generator().manual_seed(99)).long().to(clean.device)
        with acc2.accumulate(model2):
            loss = F.mse_loss(model2(sched.add_noise(clean, noise, ts), ts, return_dict=False)[0], noise)
            acc2.backward(loss)
            acc2.clip_grad_norm_(model2.parameters(), 1.0)
            opt2w.step()
            lrs2w.step()
            opt2w.zero_grad()
    got = torch.cat([p.detach().reshape(-1) for p in model2.parameters()])
    assert torch.equal(got, want)
