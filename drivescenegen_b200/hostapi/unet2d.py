"""``UNet2DModel`` with the diffusers 0.20.0 call surface, executed by libdsg_b200 on sm_100a.

Reference call sites: construction ``DriveSceneGen/scripts/train.py:39-57``; ``model(noisy, timesteps,
return_dict=False)[0]`` ``DriveSceneGen/pipeline/training_pipeline.py:84``; ``unet(image, t).sample`` inside
``DDPMPipeline.__call__`` (``DriveSceneGen/scripts/generation.py:14``); ``from_pretrained(dir, subfolder="unet")``
``DriveSceneGen/scripts/train.py:59``.

This class holds the parameter tree under upstream's names (state-dict compatible) and the
constructor/validation logic; it contains NO torch arithmetic.  ``forward`` on CUDA tensors runs ``UNetEngine``
(hand-written CUDA behind the C ABI); on CPU tensors it raises.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from .._lib import DsgError
from ..engine import UNetEngine
from .configuration import ConfigMixin

WEIGHTS_NAME = "diffusion_pytorch_model.bin"
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"


@dataclass
class UNet2DOutput:
    sample: torch.Tensor


class _Params(nn.Module):
    """Pure parameter container (children are torch.nn layers used only for their parameters and default init)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("dsg_b200 sub-modules are parameter containers; call UNet2DModel.forward")


def _resnet(cin, cout, temb, groups, eps) -> _Params:
    m = _Params()
    m.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
    m.conv1 = nn.Conv2d(cin, cout, 3, stride=1, padding=1)
    m.time_emb_proj = nn.Linear(temb, cout)
    m.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
    m.dropout = nn.Dropout(0.0)
    m.conv2 = nn.Conv2d(cout, cout, 3, stride=1, padding=1)
    m.nonlinearity = nn.SiLU()
    m.conv_shortcut = nn.Conv2d(cin, cout, 1, stride=1, padding=0) if cin != cout else None
    return m


def _attention(ch, groups, eps) -> _Params:
    m = _Params()
    m.group_norm = nn.GroupNorm(groups, ch, eps=eps, affine=True)
    m.to_q = nn.Linear(ch, ch, bias=True)
    m.to_k = nn.Linear(ch, ch, bias=True)
    m.to_v = nn.Linear(ch, ch, bias=True)
    m.to_out = nn.ModuleList([nn.Linear(ch, ch, bias=True), nn.Dropout(0.0)])
    return m


def _sampler(ch, stride) -> _Params:
    m = _Params()
    m.conv = nn.Conv2d(ch, ch, 3, stride=stride, padding=1)
    return m


_DOWN_TYPES = {"DownBlock2D": False, "AttnDownBlock2D": True}
_UP_TYPES = {"UpBlock2D": False, "AttnUpBlock2D": True}


class UNet2DModel(nn.Module, ConfigMixin):
    config_name = "config.json"

    def __init__(self, sample_size: Optional[Union[int, Tuple[int, int]]] = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, time_embedding_type: str = "positional",
                 freq_shift: int = 0, flip_sin_to_cos: bool = True,
                 down_block_types: Tuple[str, ...] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D",
                                                      "AttnDownBlock2D"),
                 up_block_types: Tuple[str, ...] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels: Tuple[int, ...] = (224, 448, 672, 896), layers_per_block: int = 2,
                 mid_block_scale_factor: float = 1, downsample_padding: int = 1, downsample_type: str = "conv",
                 upsample_type: str = "conv", act_fn: str = "silu", attention_head_dim: Optional[int] = 8,
                 norm_num_groups: int = 32, norm_eps: float = 1e-5, resnet_time_scale_shift: str = "default",
                 add_attention: bool = True, class_embed_type: Optional[str] = None,
                 num_class_embeds: Optional[int] = None):
        super().__init__()
        self.register_to_config(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            center_input_sample=center_input_sample, time_embedding_type=time_embedding_type, freq_shift=freq_shift,
            flip_sin_to_cos=flip_sin_to_cos, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, mid_block_scale_factor=mid_block_scale_factor,
            downsample_padding=downsample_padding, downsample_type=downsample_type, upsample_type=upsample_type,
            act_fn=act_fn, attention_head_dim=attention_head_dim, norm_num_groups=norm_num_groups, norm_eps=norm_eps,
            resnet_time_scale_shift=resnet_time_scale_shift, add_attention=add_attention,
            class_embed_type=class_embed_type, num_class_embeds=num_class_embeds)
        self.sample_size = sample_size
        # ---- input validation (upstream raises ValueError for these)
        if len(down_block_types) != len(up_block_types):
            raise ValueError(f"Must provide the same number of `down_block_types` as `up_block_types`. "
                             f"`down_block_types`: {down_block_types}. `up_block_types`: {up_block_types}.")
        if len(block_out_channels) != len(down_block_types):
            raise ValueError(f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                             f"`block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        for t in down_block_types:
            if t not in _DOWN_TYPES:
                raise ValueError(f"{t} does not exist.")
        for t in up_block_types:
            if t not in _UP_TYPES:
                raise ValueError(f"{t} does not exist.")
        unsupported = []
        if time_embedding_type != "positional": unsupported.append("time_embedding_type")
        if downsample_type != "conv" or upsample_type != "conv": unsupported.append("down/upsample_type")
        if downsample_padding != 1: unsupported.append("downsample_padding")
        if act_fn not in ("silu", "swish"): unsupported.append("act_fn")
        if resnet_time_scale_shift != "default": unsupported.append("resnet_time_scale_shift")
        if class_embed_type is not None or num_class_embeds is not None: unsupported.append("class embedding")
        if unsupported:
            raise NotImplementedError("dsg_b200 UNet2DModel: unsupported option(s): " + ", ".join(unsupported))

        boc = list(block_out_channels)
        temb = boc[0] * 4
        g, eps = norm_num_groups, norm_eps
        self.conv_in = nn.Conv2d(in_channels, boc[0], kernel_size=3, padding=(1, 1))
        self.time_embedding = _Params()
        self.time_embedding.linear_1 = nn.Linear(boc[0], temb)
        self.time_embedding.linear_2 = nn.Linear(temb, temb)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            blk = _Params()
            resnets, attns = [], []
            for j in range(layers_per_block):
                resnets.append(_resnet(in_ch if j == 0 else out_ch, out_ch, temb, g, eps))
                if _DOWN_TYPES[t]:
                    attns.append(_attention(out_ch, g, eps))
            if _DOWN_TYPES[t]:
                blk.attentions = nn.ModuleList(attns)
            blk.resnets = nn.ModuleList(resnets)
            blk.downsamplers = nn.ModuleList([_sampler(out_ch, 2)]) if i != len(boc) - 1 else None
            self.down_blocks.append(blk)
        mid = _Params()
        r0 = _resnet(boc[-1], boc[-1], temb, g, eps)
        at = _attention(boc[-1], g, eps) if add_attention else None
        r1 = _resnet(boc[-1], boc[-1], temb, g, eps)
        mid.attentions = nn.ModuleList([at])
        mid.resnets = nn.ModuleList([r0, r1])
        self.mid_block = mid
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            blk = _Params()
            resnets, attns = [], []
            for j in range(layers_per_block + 1):
                skip = in_ch if j == layers_per_block else out_ch
                r_in = prev if j == 0 else out_ch
                resnets.append(_resnet(r_in + skip, out_ch, temb, g, eps))
                if _UP_TYPES[t]:
                    attns.append(_attention(out_ch, g, eps))
            if _UP_TYPES[t]:
                blk.attentions = nn.ModuleList(attns)
            blk.resnets = nn.ModuleList(resnets)
            blk.upsamplers = nn.ModuleList([_sampler(out_ch, 1)]) if i != len(boc) - 1 else None
            self.up_blocks.append(blk)
        num_groups_out = norm_num_groups if norm_num_groups is not None else min(boc[0] // 4, 32)
        self.conv_norm_out = nn.GroupNorm(num_channels=boc[0], num_groups=num_groups_out, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, kernel_size=3, padding=1)
        self._engine: Optional[UNetEngine] = None
        self._engine_key = None

    # ------------------------------------------------------------------ ModelMixin surface
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    def num_parameters(self, only_trainable: bool = False) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad or not only_trainable)

    # ------------------------------------------------------------------ engine management
    def _weights_key(self):
        # _weights_epoch: bumped by optimizers that update the parameters behind autograd's back (fused AdamW kernel)
        return (getattr(self, "_weights_epoch", 0),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, train: bool = False) -> UNetEngine:
        """The CUDA engine with weights packed from the CURRENT parameter values (re-packed when they change).
        train=True also packs the data-gradient forms of the conv weights (and keeps doing so from then on)."""
        dev = self.device
        if dev.type != "cuda":
            raise DsgError("UNet2DModel is on the CPU: move it to a B200 (`.to('cuda')`); dsg_b200 has no CPU path")
        if self._engine is None or self._engine.device != dev:
            self._engine = UNetEngine(dict(self.config), dev)
            self._engine_key = None
        if train and not self._engine.train_packs:
            self._engine.train_packs = True
            self._engine_key = None
        key = self._weights_key()
        if key != self._engine_key:
            self._engine.load_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    # ------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                class_labels: Optional[torch.Tensor] = None, return_dict: bool = True):
        if not sample.is_cuda:
            raise DsgError("UNet2DModel.forward: CUDA tensors required (dsg_b200 has no CPU arithmetic path)")
        if torch.is_grad_enabled() and (sample.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .training import unet_forward_with_grad  # backward kernels (training path)
            out = unet_forward_with_grad(self, sample, timestep)
            return UNet2DOutput(sample=out) if return_dict else (out,)
        b = sample.shape[0]
        dev = sample.device
        t = timestep
        if not torch.is_tensor(t):
            t = torch.full((b,), float(t), dtype=torch.float32, device=dev)
        else:
            t = t.to(device=dev, dtype=torch.float32).reshape(-1)
            if t.numel() == 1:
                t = t.expand(b)
            t = t.contiguous()
        if t.numel() != b:
            raise ValueError("timestep must be a scalar or have one entry per sample")
        in_dtype = sample.dtype
        out = self.engine().forward(sample.float().contiguous(), t)
        if in_dtype != torch.float32:
            out = out.to(in_dtype)
        if not return_dict:
            return (out,)
        return UNet2DOutput(sample=out)

    # ------------------------------------------------------------------ (de)serialisation
    def save_pretrained(self, save_directory: str, safe_serialization: bool = False, variant: Optional[str] = None,
                        **kwargs):
        os.makedirs(save_directory, exist_ok=True)
        self.save_config(save_directory)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        name = SAFETENSORS_WEIGHTS_NAME if safe_serialization else WEIGHTS_NAME
        if variant is not None:
            stem, ext = name.rsplit(".", 1)
            name = f"{stem}.{variant}.{ext}"
        path = os.path.join(save_directory, name)
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, path, metadata={"format": "pt"})
        else:
            torch.save(sd, path)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None,
                        variant: Optional[str] = None, torch_dtype: Optional[torch.dtype] = None, **kwargs):
        d = pretrained_model_name_or_path
        if subfolder:
            d = os.path.join(d, subfolder)
        if not os.path.isdir(d):
            raise EnvironmentError(f"{d} is not a local directory (dsg_b200 loads local checkpoints only)")
        cfg = cls.load_config(d)
        model = cls(**cfg)
        cands = []
        for base in (SAFETENSORS_WEIGHTS_NAME, WEIGHTS_NAME):
            stem, ext = base.rsplit(".", 1)
            cands.append(f"{stem}.{variant}.{ext}" if variant else base)
        path = next((os.path.join(d, c) for c in cands if os.path.isfile(os.path.join(d, c))), None)
        if path is None:
            raise EnvironmentError(f"Error no file named {cands[1]} found in directory {d}.")
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            sd = torch.load(path, map_location="cpu", weights_only=True)
        # weights are up-cast to the module dtype (fp32) unless torch_dtype is given — upstream behaviour
        model.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()}, strict=True)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        model.eval()
        return model
