"""oracle/ — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch, CPU-runnable restatement of the arithmetic the reference borrows from
``diffusers==0.20.0`` (``requirements.txt:15``): ``UNet2DModel``, ``DDPMScheduler``,
``DDIMScheduler`` and the ``DDPMPipeline`` sampling loop, as called from
``DriveSceneGen/scripts/train.py:39-71``, ``DriveSceneGen/pipeline/training_pipeline.py:26-107``
and ``DriveSceneGen/scripts/generation.py:7-20``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product (``drivescenegen_b200``) never does: its hot path is the
sm_100a CUDA library and it fails loudly when that library is missing.

PARITY STATUS
-------------
* Schedulers: PINNED.  The upstream known-answer loops (``test_scheduler_ddpm.py::test_full_loop_no_noise``
  -> 258.9606 / 0.3372 and ``test_scheduler_ddim.py::test_full_loop_no_noise`` -> 172.0067 / 0.223967)
  are reproduced by ``tests/test_oracle_kat.py``.
* U-Net: PARITY UNPINNED.  ``diffusers`` is not installed here, is not vendored in ``/root/reference`` and
  the reference ships no tests or golden vectors.  The restatement is pinned only by the upstream
  parameter counts (113,673,219 / 56,574,595 / 3,660,803), the state-dict key table and op-level
  composition from ``torch.nn.functional``.  To pin it: ``python tests/golden/make_unet_golden.py`` where
  ``diffusers==0.20.0`` is installed writes ``tests/golden/unet_golden.npz`` (eps slices of real diffusers models
  with deterministic weights); ``tests/test_oracle_unet_golden.py`` then checks this package against it.
"""
