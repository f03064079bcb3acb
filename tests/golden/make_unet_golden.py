#!/usr/bin/env python
"""Pin the U-Net oracle against REAL diffusers — one command, to be run wherever ``diffusers==0.20.0`` is importable
(it is not in the build container: no index access, not vendored in /root/reference, so ``oracle/unet.py`` is
"parity unpinned" until this has been run once):

    pip install diffusers==0.20.0 && python tests/golden/make_unet_golden.py

Writes tests/golden/unet_golden.npz; ``tests/test_oracle_unet_golden.py`` then checks ``oracle.unet.OracleUNet2D``
against it on CPU (and skips, saying so, while the file is absent).

The weights are NOT stored (56.6 M parameters): both sides fill every state-dict tensor, in sorted key order, from one
``torch.Generator`` stream (``fill_state_dict`` below — plain torch, no diffusers code), so the fixture holds only the
inputs' seed, a slice of the predicted noise per case and its checksums.  Cases: the reference's model
(DriveSceneGen/scripts/train.py:39-57) at 64 x 64 and 96 x 160, and the attention variant at 32 x 32.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = dict(in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(64, 128, 256, 512),
           down_block_types=("DownBlock2D",) * 4, up_block_types=("UpBlock2D",) * 4)
ATTN = dict(in_channels=3, out_channels=3, layers_per_block=1, block_out_channels=(64, 128),
            down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"))
CASES = [("ref64", REF, (2, 3, 64, 64), [999.0, 3.0]),
         ("ref96x160", REF, (1, 3, 96, 160), [500.0]),
         ("attn32", ATTN, (2, 3, 32, 32), [10.0, 750.0])]


def fill_state_dict(module: torch.nn.Module, seed: int) -> None:
    """Deterministic, framework-independent weights: tensors visited in sorted key order; weights of >= 2 dims
    ~ N(0, 1 / fan_in), norm gains 1 + 0.1 N, everything else 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for k in sorted(sd):
        t = sd[k]
        if not t.is_floating_point():
            continue
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if t.dim() >= 2:
            r = r / float(t[0].numel()) ** 0.5
        elif k.endswith("weight"):
            r = 1.0 + 0.1 * r
        else:
            r = 0.1 * r
        t.copy_(r)


def inputs(shape, seed: int) -> torch.Tensor:
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def summarize(eps: torch.Tensor) -> dict:
    flat = eps.double().flatten()
    idx = torch.linspace(0, flat.numel() - 1, 256).long()
    return {"slice": flat[idx].numpy(), "sum": float(flat.sum()), "abs_sum": float(flat.abs().sum()),
            "shape": np.array(eps.shape)}


def main():
    import diffusers
    from diffusers import UNet2DModel
    assert diffusers.__version__.startswith("0.20"), f"the reference pins diffusers==0.20.0, found {diffusers.__version__}"
    torch.set_grad_enabled(False)
    out = {"diffusers_version": np.array(diffusers.__version__)}
    for name, cfg, shape, ts in CASES:
        net = UNet2DModel(sample_size=shape[2:], **cfg).eval()
        fill_state_dict(net, 20261017)
        x = inputs(shape, 7)
        t = torch.tensor((ts * shape[0])[: shape[0]])
        eps = net(x, t).sample
        for k, v in summarize(eps).items():
            out[f"{name}/{k}"] = np.asarray(v)
        out[f"{name}/params"] = np.array(sum(p.numel() for p in net.parameters()))
    path = os.path.join(HERE, "unet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
