"""TEST-ONLY CPU plumbing: lets the test-suite drive the reference's unmodified scripts through the `diffusers` /
`accelerate` shim packages in a container without a GPU (BASELINE.json configs[0]).

The product (`drivescenegen_b200`) has no CPU arithmetic path and no hook for one: every host-API entry point raises
`DsgError` on CPU tensors.  `install(monkeypatch)` patches the host-API classes *from the outside*, for the duration of
one test, so that CPU tensors are served by the CPU oracle (`oracle/`); CUDA tensors still take the original methods.
"""
import torch


def install(monkeypatch):
    from drivescenegen_b200.hostapi import pipeline as pl
    from drivescenegen_b200.hostapi import schedulers as sc
    from drivescenegen_b200.hostapi import unet2d as un
    from oracle.schedulers import OracleDDPMScheduler
    from oracle.unet import OracleUNet2D

    nets = {}
    orig_forward = un.UNet2DModel.forward
    orig_add_noise = sc._SchedulerBase.add_noise
    orig_step = sc.DDPMScheduler.step
    orig_call = pl.DDPMPipeline.__call__

    def forward(self, sample, timestep, class_labels=None, return_dict=True):
        if sample.is_cuda:
            return orig_forward(self, sample, timestep, class_labels, return_dict)
        if id(self) not in nets:
            cfg = {k: self.config[k] for k in ("sample_size", "in_channels", "out_channels", "down_block_types",
                                               "up_block_types", "block_out_channels", "layers_per_block",
                                               "attention_head_dim", "norm_num_groups", "norm_eps", "add_attention")}
            nets[id(self)] = OracleUNet2D(**cfg)
        out = torch.func.functional_call(nets[id(self)], dict(self.named_parameters()), (sample, timestep))[0]
        return un.UNet2DOutput(sample=out) if return_dict else (out,)

    def add_noise(self, original_samples, noise, timesteps):
        if original_samples.is_cuda:
            return orig_add_noise(self, original_samples, noise, timesteps)
        return OracleDDPMScheduler().add_noise(original_samples, noise, timesteps)

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, variance_noise=None):
        if sample.is_cuda:
            return orig_step(self, model_output, timestep, sample, generator, return_dict, variance_noise)
        o = OracleDDPMScheduler()
        if self.num_inference_steps:
            o.set_timesteps(self.num_inference_steps)
        prev = o.step(model_output, int(timestep), sample, generator=generator)
        return sc.SchedulerOutput(prev_sample=prev) if return_dict else (prev,)

    @torch.no_grad()
    def call(self, batch_size=1, generator=None, num_inference_steps=1000, output_type="pil", return_dict=True):
        if self.device.type == "cuda":
            return orig_call(self, batch_size, generator, num_inference_steps, output_type, return_dict)
        image = sc.randn_tensor(self._image_shape(batch_size), generator=generator, device=self.device)
        self.scheduler.set_timesteps(num_inference_steps)
        for t in self.progress_bar(self.scheduler.timesteps):
            model_output = self.unet(image, t).sample
            image = self.scheduler.step(model_output, t, image, generator=generator).prev_sample
        image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
        if output_type == "pil":
            image = pl.numpy_to_pil(image)
        return pl.ImagePipelineOutput(images=image) if return_dict else (image,)

    monkeypatch.setattr(un.UNet2DModel, "forward", forward)
    monkeypatch.setattr(sc._SchedulerBase, "add_noise", add_noise)
    monkeypatch.setattr(sc.DDPMScheduler, "step", step)
    monkeypatch.setattr(pl.DDPMPipeline, "__call__", call)
