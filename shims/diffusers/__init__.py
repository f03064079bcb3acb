"""Drop-in `diffusers` import surface for SS47816/DriveSceneGen, served by drivescenegen_b200.

Put `<repo>/shims` on PYTHONPATH and the reference's scripts (`DriveSceneGen/scripts/train.py`,
`scripts/generation.py`, `pipeline/training_pipeline.py`) import these symbols unmodified.
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from drivescenegen_b200.hostapi import (  # noqa: E402,F401
    DDIMPipeline, DDIMScheduler, DDPMPipeline, DDPMScheduler, ImagePipelineOutput, UNet2DModel, UNet2DOutput)
from drivescenegen_b200.hostapi.configuration import DIFFUSERS_VERSION as __version__  # noqa: E402,F401
from . import optimization  # noqa: E402,F401
