"""Host-side mirror of the reference-facing interface (the `diffusers` / `accelerate` symbols the reference imports)."""
from .unet2d import UNet2DModel, UNet2DOutput  # noqa: F401
from .schedulers import DDPMScheduler, DDIMScheduler, SchedulerOutput, randn_tensor  # noqa: F401
from .pipeline import DDPMPipeline, DDIMPipeline, DenoiseSession, ImagePipelineOutput  # noqa: F401
from .optimization import get_cosine_schedule_with_warmup  # noqa: F401
from .accelerator import Accelerator, notebook_launcher  # noqa: F401
from . import raster  # noqa: F401
from .raster import RasterDataset  # noqa: F401
