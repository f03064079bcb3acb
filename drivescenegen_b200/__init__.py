"""drivescenegen_b200 — B200-native (sm_100a) denoising engine behind the `diffusers` call surface that
SS47816/DriveSceneGen uses: UNet2DModel.forward + DDPM/DDIM scheduler.step, hand-written CUDA behind a C ABI
(include/dsg_b200.h).  The host API lives in `drivescenegen_b200.hostapi`; `shims/` re-exports it as `diffusers` /
`accelerate`."""
__version__ = "0.1.0"
