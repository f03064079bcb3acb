"""CPU oracle: ``diffusers==0.20.0`` ``UNet2DModel`` restated with plain torch.nn / torch.nn.functional.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  PARITY UNPINNED for the U-Net (no upstream source,
tests or golden vectors on disk); pinned by parameter counts and state-dict keys only.

Follows (upstream, restated in /root/repo/SURVEY.md App. A; call sites in the reference):
  * constructor kwargs           — DriveSceneGen/scripts/train.py:39-57
  * forward(sample, timestep)    — DriveSceneGen/pipeline/training_pipeline.py:84,
                                   DDPMPipeline.__call__ via DriveSceneGen/scripts/generation.py:14
  * upstream modules restated: models/unet_2d.py, models/unet_2d_blocks.py (DownBlock2D, AttnDownBlock2D,
    UNetMidBlock2D, UpBlock2D, AttnUpBlock2D), models/resnet.py (ResnetBlock2D, Downsample2D, Upsample2D),
    models/attention_processor.py (Attention + AttnProcessor2_0), models/embeddings.py.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True,
                       freq_shift: float = 0.0, max_period: int = 10000) -> torch.Tensor:
    """upstream ``get_timestep_embedding`` (models/embeddings.py), scale = 1."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int, eps: float,
                 output_scale_factor: float = 1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.conv_shortcut = None
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0)
        self.output_scale_factor = output_scale_factor

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))  # dropout p=0
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class _ConvHolder(nn.Module):
    """``Downsample2D`` / ``Upsample2D`` with ``use_conv=True``: a single ``.conv`` child."""

    def __init__(self, channels: int, stride: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=stride, padding=1)


class Downsample2D(_ConvHolder):
    def __init__(self, channels):
        super().__init__(channels, 2)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(_ConvHolder):
    def __init__(self, channels):
        super().__init__(channels, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Attention(nn.Module):
    """upstream ``Attention(..., residual_connection=True, bias=True, _from_deprecated_attn_block=True)``
    evaluated by ``AttnProcessor2_0`` (torch>=2.0)."""

    def __init__(self, channels: int, head_dim: int, groups: Optional[int], eps: float,
                 rescale_output_factor: float = 1.0):
        super().__init__()
        self.heads = channels // head_dim
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True) if groups is not None else None
        self.to_q = nn.Linear(channels, channels, bias=True)
        self.to_k = nn.Linear(channels, channels, bias=True)
        self.to_v = nn.Linear(channels, channels, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])
        self.rescale_output_factor = rescale_output_factor

    def forward(self, x):
        res = x
        b, c, h, w = x.shape
        hs = x.view(b, c, h * w).transpose(1, 2)
        if self.group_norm is not None:
            hs = self.group_norm(hs.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(hs), self.to_k(hs), self.to_v(hs)
        d = c // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, c)
        o = self.to_out[1](self.to_out[0](o))
        o = o.transpose(-1, -2).reshape(b, c, h, w)
        return (o + res) / self.rescale_output_factor


class DownBlock(nn.Module):
    """``DownBlock2D`` (attn=False) / ``AttnDownBlock2D`` (attn=True)."""

    def __init__(self, in_ch, out_ch, temb_ch, num_layers, groups, eps, add_downsample, attn, head_dim):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, groups, eps))
            if attn:
                attentions.append(Attention(out_ch, head_dim, groups, eps))
        if attn:
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None
        self.has_attn = attn

    def forward(self, x, temb):
        outs = ()
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class MidBlock(nn.Module):
    """``UNetMidBlock2D`` with ``num_layers=1``: resnet, attention, resnet."""

    def __init__(self, ch, temb_ch, groups, eps, add_attention, head_dim, output_scale_factor=1.0):
        super().__init__()
        resnets = [ResnetBlock2D(ch, ch, temb_ch, groups, eps, output_scale_factor)]
        attentions = []
        if add_attention:
            attentions.append(Attention(ch, head_dim, groups, eps, output_scale_factor))
        else:
            attentions.append(None)
        resnets.append(ResnetBlock2D(ch, ch, temb_ch, groups, eps, output_scale_factor))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)

    def forward(self, x, temb):
        x = self.resnets[0](x, temb)
        for attn, r in zip(self.attentions, self.resnets[1:]):
            if attn is not None:
                x = attn(x)
            x = r(x, temb)
        return x


class UpBlock(nn.Module):
    """``UpBlock2D`` (attn=False) / ``AttnUpBlock2D`` (attn=True)."""

    def __init__(self, in_ch, prev_ch, out_ch, temb_ch, num_layers, groups, eps, add_upsample, attn, head_dim):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            skip_ch = in_ch if i == num_layers - 1 else out_ch
            r_in = prev_ch if i == 0 else out_ch
            resnets.append(ResnetBlock2D(r_in + skip_ch, out_ch, temb_ch, groups, eps))
            if attn:
                attentions.append(Attention(out_ch, head_dim, groups, eps))
        if attn:
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None
        self.has_attn = attn

    def forward(self, x, skips: Tuple[torch.Tensor, ...], temb):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips[-1]], dim=1)
            skips = skips[:-1]
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


_DOWN = {"DownBlock2D": False, "AttnDownBlock2D": True}
_UP = {"UpBlock2D": False, "AttnUpBlock2D": True}


class OracleUNet2D(nn.Module):
    """Restatement of ``diffusers.UNet2DModel`` (0.20.0) for positional time embedding, conv down/up-sampling,
    ``resnet_time_scale_shift="default"``, no class embedding — everything the reference exercises."""

    def __init__(self, sample_size: Union[int, Tuple[int, int], None] = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, freq_shift: int = 0,
                 flip_sin_to_cos: bool = True,
                 down_block_types: Sequence[str] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types: Sequence[str] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels: Sequence[int] = (224, 448, 672, 896), layers_per_block: int = 2,
                 mid_block_scale_factor: float = 1.0, attention_head_dim: Optional[int] = 8,
                 norm_num_groups: int = 32, norm_eps: float = 1e-5, add_attention: bool = True):
        super().__init__()
        if len(down_block_types) != len(up_block_types) or len(block_out_channels) != len(down_block_types):
            raise ValueError("down_block_types, up_block_types and block_out_channels must have the same length")
        for t in down_block_types:
            if t not in _DOWN:
                raise ValueError(f"{t} does not exist.")
        for t in up_block_types:
            if t not in _UP:
                raise ValueError(f"{t} does not exist.")
        self.sample_size = sample_size
        self.in_channels = in_channels
        self.center_input_sample = center_input_sample
        self.flip_sin_to_cos, self.freq_shift = flip_sin_to_cos, freq_shift
        boc = list(block_out_channels)
        temb_ch = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            hd = attention_head_dim if attention_head_dim is not None else out_ch
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb_ch, layers_per_block, norm_num_groups, norm_eps,
                                              i != len(boc) - 1, _DOWN[t], hd))
        hd = attention_head_dim if attention_head_dim is not None else boc[-1]
        self.mid_block = MidBlock(boc[-1], temb_ch, norm_num_groups, norm_eps, add_attention, hd,
                                  mid_block_scale_factor)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            hd = attention_head_dim if attention_head_dim is not None else out_ch
            self.up_blocks.append(UpBlock(in_ch, prev, out_ch, temb_ch, layers_per_block + 1, norm_num_groups,
                                          norm_eps, i != len(boc) - 1, _UP[t], hd))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.time_dim = boc[0]

    def forward(self, sample: torch.Tensor, timestep, return_dict: bool = False):
        if self.center_input_sample:
            sample = 2 * sample - 1.0
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        t_emb = timestep_embedding(t, self.time_dim, self.flip_sin_to_cos, self.freq_shift)
        emb = self.time_embedding(t_emb.to(self.conv_in.weight.dtype))
        x = self.conv_in(sample)
        skips = (x,)
        for blk in self.down_blocks:
            x, outs = blk(x, emb)
            skips += outs
        x = self.mid_block(x, emb)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            res, skips = skips[-n:], skips[:-n]
            x = blk(x, res, emb)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x,)
