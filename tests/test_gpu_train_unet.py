"""GPU parity of the whole training path: UNet2DModel forward + hand-written backward vs the CPU oracle under torch
autograd (fp32).  Tolerance (fp16 operands / activations / activation gradients, fp32 accumulation): relative L2 error of
each parameter gradient <= 4e-2, of all gradients together <= 2e-2, of the output <= 1e-2."""
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


@pytest.mark.parametrize("name,cfg,size,batch", [("c1", CFG_C1, 64, 2), ("ref", CFG_REF, 64, 2), ("ref128", CFG_REF, 128, 1)])
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
    assert _rel(out.detach().cpu(), ref_out.detach()) < 1e-2
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
    assert (num / den) ** 0.5 < 2e-2, ((num / den) ** 0.5, worst[:5])
    assert worst[0][0] < 4e-2, worst[:8]


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
