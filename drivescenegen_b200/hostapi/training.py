"""Training path (forward with autograd) — NOT BUILT YET in this round.

``UNet2DModel.forward`` under ``torch.enable_grad()`` (DriveSceneGen/pipeline/training_pipeline.py:84-86) needs the
dgrad / wgrad / GroupNorm-backward / attention-backward kernels.  Until they exist this raises
loudly instead of silently differentiating through a PyTorch re-implementation.
"""


def unet_forward_with_grad(model, sample, timestep):
    raise NotImplementedError(
        "dsg_b200: the backward kernels (training path) are not built yet; run inference under torch.no_grad()")
