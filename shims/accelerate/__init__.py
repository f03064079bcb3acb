"""Drop-in `accelerate` import surface (Accelerator, notebook_launcher) served by drivescenegen_b200."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from drivescenegen_b200.hostapi.accelerator import Accelerator, notebook_launcher  # noqa: E402,F401

__version__ = "0.22.0"
