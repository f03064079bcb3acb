"""Test hook: lets the TEST SUITE plug a CPU implementation behind the host API for plumbing tests.

The product has no CPU arithmetic path.  ``tests/`` may register the CPU oracle here so that the reference's
unmodified scripts (BASELINE.json configs[0]: 64x64, 2 blocks, 1 DDPM step, CPU) can be driven through the shim
packages without a GPU.  Nothing in ``drivescenegen_b200`` ever registers a backend by itself, and a registered
backend is only consulted for CPU tensors — CUDA tensors always take the libdsg_b200 path.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

_CPU_BACKEND: Dict[str, Callable] = {}


def register_cpu_backend(name: str, fn: Optional[Callable]) -> None:
    """name in {"unet_forward", "ddpm_step", "ddim_step", "add_noise"}; fn=None removes it."""
    if fn is None:
        _CPU_BACKEND.pop(name, None)
    else:
        _CPU_BACKEND[name] = fn


def cpu_backend(name: str) -> Optional[Callable]:
    return _CPU_BACKEND.get(name)


def clear_cpu_backends() -> None:
    _CPU_BACKEND.clear()
