#!/usr/bin/env python
"""bench.py — denoise-steps/s on 256x256x3 rasters (BASELINE.json configs[1]) and the U-Net conv roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--size S]
                    [--scheduler ddim --size 512 --batch 8]   (configs[3])      [--workload train]   (configs[2] alone)

One "step" = one DDPM denoising step (U-Net forward + scheduler.step) over a batch of B=16 synthetic 256x256x3
samples with random-init weights of the reference architecture (DriveSceneGen/scripts/train.py:39-57).
  value       sample-steps/s (= B*K/time), inputs resident in HBM, CUDA events, max over ranks; N>1 = N independent
              replicas (sampling never communicates: SURVEY.md §8e), weak scaling.  One CUDA-graph launch per step.
  e2e         same metric through DenoiseSession.run_from_host: every step copies that step's variance noise from
              pinned host memory to the device and reads the new sample back to pinned host memory; the copies run
              on their own streams beside the next step's compute (serial copy->step->copy->sync time also reported).
  e2e_pipeline  the reference-facing call itself, DDPMPipeline.__call__ (training_pipeline.py:26-32, generation.py:14),
              K steps, numpy out: with generator=None and with a CPU generator (upstream RNG rule: host draws).
  roofline    the tcgen05 implicit-GEMM conv kernels: algorithmic conv/linear FLOPs of one step (reference op count)
              / summed CUDA-event duration of those launches in an eager, per-launch-timed replay of the same step
              (per-launch median of 5 replays); `frac_executed` counts the MACs really executed (sub-pixel upsample
              convs run 4/9 of the reference's); `traffic` = DRAM bytes per launch from an `ncu --set full` capture of the
              commit named in `traffic_source` (tools/conv_traffic.py).
  kernels     per-kernel-class records (conv, conv_cout64, conv_upsample, gn_apply, conv_in/out, attention, sched_step):
              algorithmic work, executed work, event time, fraction of the measured peak.
  whole_step  U-Net forward FLOPs / ms_per_step against the tensor peak (GroupNorm, attention, scheduler included).
  train       configs[2] in the same line: 256x256x3 training step (fwd + bwd + clip + AdamW), batch 32 per GPU, one
              flat-buffer gradient all-reduce per step when N > 1 (overlapped with the backward), with its own roofline
              (weight-gradient kernel), per-class records and whole-step fraction.
  cpu_baseline / --impl reference: the CPU oracle (plain PyTorch fp32 restatement of the reference path) on this
              box's host cores, bounded sample (batch 2).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

REF_CFG = dict(in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(64, 128, 256, 512),
               down_block_types=("DownBlock2D",) * 4, up_block_types=("UpBlock2D",) * 4)
METRIC = "denoise-steps/sec (256x256x3 raster)"
UNIT = "sample-steps/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_sustained": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "25", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            sm.sort()
            # median over the samples taken under load (upper half: idle samples before/after the region are low)
            busy = sm[len(sm) // 2:]
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are meant to use every core the box gives us."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if torch.get_num_threads() != n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def median_table(tables):
    """per-launch median over several event-timed eager replays of the same op list: one replay is exposed to single
    multi-millisecond outliers (a host hiccup between two launches lands in one launch's event pair)."""
    out = []
    for rows in zip(*tables):
        ms = sorted(r[2] for r in rows)
        out.append((rows[0][0], rows[0][1], ms[len(ms) // 2]))
    return out


def oracle_steps_per_s(batch: int, size: int, steps: int, warmup: int):
    """CPU oracle: U-Net forward + DDPM step, fp32, all host threads."""
    use_all_host_threads()
    from oracle.schedulers import OracleDDPMScheduler
    from oracle.unet import OracleUNet2D
    torch.manual_seed(0)
    net = OracleUNet2D(sample_size=(size, size), **REF_CFG).eval()
    sch = OracleDDPMScheduler()
    sch.set_timesteps(1000)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 3, size, size, generator=g)
    ts = [int(t) for t in sch.timesteps[: warmup + steps]]
    with torch.no_grad():
        for t in ts[:warmup]:
            x = sch.step(net(x, t)[0], t, x, generator=g)
        t0 = time.perf_counter()
        for t in ts[warmup:]:
            x = sch.step(net(x, t)[0], t, x, generator=g)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    # keep the whole run within a few minutes whatever K the driver passes (~1-2 s per sample-step on 8 threads)
    steps = min(steps, 40)
    v, dt = oracle_steps_per_s(batch, args.size, steps, warmup)
    cores = torch.get_num_threads()
    sample = f"batch {batch} x {steps} steps of {args.size}x{args.size}x3 (oracle port, fp32, {cores} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"256x256x3 BEV raster, full U-Net (56.6M params), DDPM sampling; CPU sample: {sample}"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def oracle_train_steps_per_s(batch: int, size: int, steps: int):
    """CPU oracle: forward + autograd backward + torch AdamW, fp32, all host threads."""
    import torch.nn.functional as F
    from oracle.unet import OracleUNet2D
    use_all_host_threads()
    torch.manual_seed(0)
    net = OracleUNet2D(sample_size=(size, size), **REF_CFG).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-5)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(batch, 3, size, size, generator=g)
    noise = torch.randn(batch, 3, size, size, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    dt = 0.0
    for i in range(steps + 1):
        t0 = time.perf_counter()
        loss = F.mse_loss(net(x, t)[0], noise)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        if i > 0:
            dt += time.perf_counter() - t0
    return batch * steps / dt, dt


def train_record(args, dev, rank, world, K, W, with_cpu_baseline):
    """BASELINE configs[2]: 256x256x3 training step (fwd + bwd + clip + AdamW), batch 32 per GPU, one gradient
    all-reduce per step when N > 1.  Runs on every rank; returns the record (rank 0) or None."""
    import torch.nn.functional as F
    from drivescenegen_b200 import _lib
    from drivescenegen_b200.hostapi import Accelerator, DDPMScheduler, UNet2DModel, get_cosine_schedule_with_warmup
    B, S = args.train_batch, args.size
    acc = Accelerator(mixed_precision="fp16", gradient_accumulation_steps=1)   # initialises NCCL when world > 1
    torch.manual_seed(0)
    model = UNet2DModel(sample_size=(S, S), **REF_CFG).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
    lr_sched = get_cosine_schedule_with_warmup(optimizer=opt, num_warmup_steps=500, num_training_steps=100000)
    model, opt, lr_sched = acc.prepare(model, opt, lr_sched)
    sched = DDPMScheduler()
    g = torch.Generator().manual_seed(7 + rank)
    x_host = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory()
    x = x_host.to(dev)
    noise = torch.randn(B, 3, S, S, generator=g).to(dev)
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(8 + rank)).to(dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(from_host: bool):
        xi = x_host.to(dev, non_blocking=True) if from_host else x
        noisy = sched.add_noise(xi, noise, t)
        with acc.accumulate(model):
            pred = model(noisy, t, return_dict=False)[0]
            loss = F.mse_loss(pred, noise)
            acc.backward(loss)
            acc.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            lr_sched.step()
            opt.zero_grad()
        if from_host:
            loss_host.copy_(loss.detach(), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step(False)
    barrier()
    sampler = ClockSampler(dev.index) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()   # kernels of the K timed steps (graph replays report their kernel count to the library)
    e0.record()
    for _ in range(K):
        step(False)
    e1.record()
    timed_launches = _lib.launch_count() - n0
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step(True)
    barrier()
    e2e_s = time.perf_counter() - t0
    # the collective alone (same buffer, same stream), to state how much of the step it is
    ar_ms = None
    if world > 1:
        import torch.distributed as dist
        fg = getattr(model, "_flat_grads", None)
        if fg is not None:
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(5):
                acc._reduce_mean(fg.flat[0])
            a1.record()
            barrier()
            ar_ms = a0.elapsed_time(a1) / 5
        tt = torch.tensor([ms, e2e_s * 1000.0, ar_ms or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s, ar_ms = tt[0].item(), tt[1].item() / 1000.0, tt[2].item()
    flat_p = getattr(model, "_flat_params", None)
    checksum = float(flat_p.double().sum().item()) if flat_p is not None else None   # same on every rank after N steps
    if rank != 0:
        return None
    pk = peaks()
    # per-launch table of one forward + backward (eager, CUDA events)
    eng = model.engine(train=True)
    prog = next(iter(eng.train_programs.values()))
    noisy = sched.add_noise(x, noise, t)
    tf = t.float()
    dout = torch.randn(B, 3, S, S, device=dev) * 1e-3
    prog.run(noisy, tf)
    prog.backward_timed(dout, noisy)
    fwd = median_table([prog.run_timed(noisy, tf) for _ in range(3)])
    bwd = median_table([prog.backward_timed(dout, noisy) for _ in range(3)])

    def agg(table, name):
        sel = [(meta, m) for n, meta, m in table if n == name]
        return sum(m for _, m in sel), sum((meta.get("flops") or 0) for meta, _ in sel), len(sel)
    wg_ms, wg_fl, wg_n = agg(bwd, "wgrad")
    dg_ms, dg_fl, dg_n = agg(bwd, "dgrad")
    cv_ms, cv_fl, cv_n = agg(fwd, "conv")
    gnb_ms, _, gnb_n = agg(bwd, "gn_bwd")
    gnb_bytes = sum(meta.get("bytes", 0) for n, meta, m in bwd if n == "gn_bwd")
    gnb_alg = sum(meta.get("bytes_alg", 0) for n, meta, m in bwd if n == "gn_bwd")
    fwd_ms, bwd_ms = sum(m for _, _, m in fwd), sum(m for _, _, m in bwd)
    step_flops = 3 * sum((meta.get("flops") or 0) for _, meta, _ in fwd)
    roofline = {"bound": "tensor", "kernel": "wgrad_kernel (all conv / linear weight-gradient launches of one step)",
                "achieved": wg_fl / (wg_ms * 1e-3) / 1e12, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                "frac": wg_fl / (wg_ms * 1e-3) / 1e12 / pk["tflops_sustained"],
                "peak_source": pk["source"] + " bf16 sustained (cuBLAS)", "traffic": None, "launches": wg_n,
                "avg_launch_ms": wg_ms / max(wg_n, 1), "algorithmic_flops_per_step": wg_fl,
                "share_of_step": wg_ms / (fwd_ms + bwd_ms)}
    kernels = {
        "wgrad": {"bound": "tensor", "launches": wg_n, "ms": wg_ms, "achieved_tflops": wg_fl / (wg_ms * 1e-3) / 1e12,
                  "frac": wg_fl / (wg_ms * 1e-3) / 1e12 / pk["tflops_sustained"]},
        "dgrad": {"bound": "tensor", "launches": dg_n, "ms": dg_ms, "achieved_tflops": dg_fl / (dg_ms * 1e-3) / 1e12,
                  "frac": dg_fl / (dg_ms * 1e-3) / 1e12 / pk["tflops_sustained"]},
        "fwd_conv": {"bound": "tensor", "launches": cv_n, "ms": cv_ms, "achieved_tflops": cv_fl / (cv_ms * 1e-3) / 1e12,
                     "frac": cv_fl / (cv_ms * 1e-3) / 1e12 / pk["tflops_sustained"]},
        # algorithmic bytes of GroupNorm+SiLU backward: read x, read dy, write dx (3 passes of the tensor); executed =
        # what the two kernels actually move (statistics pass + apply pass + the shortcut / skip addends)
        "gn_bwd": {"bound": "hbm", "launches": gnb_n, "ms": gnb_ms,
                   "achieved_gbs_algorithmic": gnb_alg / (gnb_ms * 1e-3) / 1e9,
                   "frac": gnb_alg / (gnb_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                   "achieved_gbs_executed": gnb_bytes / (gnb_ms * 1e-3) / 1e9,
                   "frac_executed": gnb_bytes / (gnb_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                   "note": "reported against HBM; ncu shows the two kernels instruction-issue-bound (52-61 % issue-active "
                           "at 23 % occupancy, profiles/r2o_ncu_full_gn_bwd_streamed_vs_register.txt)"},
        "attention_bwd": {"bound": "mufu", "ms": agg(bwd, "attention_bwd")[0]},
    }
    breakdown = {"fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "fwd_conv_ms": cv_ms, "wgrad_ms": wg_ms, "dgrad_ms": dg_ms,
                 "gn_bwd_ms": gnb_ms, "attention_bwd_ms": agg(bwd, "attention_bwd")[0],
                 "other_bwd_ms": bwd_ms - wg_ms - dg_ms - gnb_ms - agg(bwd, "attention_bwd")[0]}
    if args.profile_out:
        path = args.profile_out if args.workload == "train" else args.profile_out.replace(".json", "_train.json")
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as f:
            json.dump({"batch": B, "size": S,
                       "forward": [{"op": n, **meta, "ms": m} for n, meta, m in fwd],
                       "backward": [{"op": n, **meta, "ms": m} for n, meta, m in bwd]}, f, indent=1)
    cpu = None
    if with_cpu_baseline:
        v, dt = oracle_train_steps_per_s(2, S, 1)
        cores = torch.get_num_threads()
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"batch 2 x 1 train step of {S}x{S}x3 after 1 warm-up (oracle, fp32 autograd + AdamW, {cores} threads, {dt:.1f} s)"}
    return {"metric": "train samples/sec (256x256x3, fwd+bwd+clip+AdamW)", "value": world * B * K / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands / activations / activation gradients, fp32 accumulate, fp32 master weights + AdamW",
            "data": "synthetic",
            "config": {"workload": f"{S}x{S}x3 training step (U-Net fwd + bwd + grad clip + AdamW, GradScaler), batch {B} "
                                   f"per GPU, random-init reference U-Net (56.6M params)",
                       "batch_per_gpu": B, "parallelism": f"dp{world} (one NCCL all-reduce over the flat fp32 gradient "
                                                          f"buffer per step)" if world > 1 else "single GPU",
                       "l2": "per-step working set (tens of GB of activations) >> 126 MB L2, no flush needed"},
            "clocks": clocks,
            "e2e": {"value": world * B * K / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": 1000.0 * e2e_s / K},
            "allreduce": None if world == 1 else {
                "bytes": 4 * sum(p.numel() for p in model.parameters()), "ms_alone": ar_ms,
                "overlap": os.environ.get("DSG_AR_OVERLAP", "1") != "0",
                "how": "one flat fp32 gradient buffer, averaged in place by NCCL; its tail (up / mid / output layers, "
                       "64 % of the bytes) is issued from inside the backward pass and runs under the down path's "
                       "backward, the head follows when the backward returns"},
            "param_checksum": checksum,
            "gpu_launches": timed_launches, "roofline": roofline, "kernels": kernels, "breakdown": breakdown,
            "whole_step": {"flops": step_flops, "ms": ms / K,
                           "frac": step_flops / (ms / K * 1e-3) / 1e12 / pk["tflops_sustained"]},
            "cpu_baseline": cpu, "train_flops_per_step": step_flops}


def run_train(args):
    """`--workload train`: the configs[2] record as its own JSON line."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    S = args.size
    W, K = max(3, args.warmup), max(1, args.steps)
    if args.impl == "reference":
        if rank != 0:
            return
        v, dt = oracle_train_steps_per_s(2, S, max(1, min(K, 3)))
        cores = torch.get_num_threads()
        sample = f"batch 2 x {max(1, min(K, 3))} train steps of {S}x{S}x3 (oracle port, fp32 autograd + AdamW, {cores} threads)"
        print(json.dumps({"impl": "reference", "metric": "train samples/sec (256x256x3)", "value": v,
                          "unit": "samples/s", "n_gpus": args.gpus, "higher_is_better": True,
                          "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return
    assert torch.cuda.is_available(), "bench.py needs a GPU (the product has no CPU path)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    rec = train_record(args, dev, rank, world, K, W, world == 1 and not args.no_cpu_baseline)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def time_sched_step(sess, reps=20):
    """the scheduler-step kernel alone (3 reads + 1 write of the sample), eager, CUDA events."""
    from drivescenegen_b200._lib import check
    lib = sess.lib
    fn = lib.dsg_ddim_step if sess.ddim else lib.dsg_ddpm_step
    st = torch.cuda.current_stream(sess.dev).cuda_stream
    row = torch.full((1,), 500, dtype=torch.int32, device=sess.dev)
    a = (sess.eps.data_ptr(), sess.xb[0].data_ptr(), sess.zb[0].data_ptr(), sess.xb[1].data_ptr(), sess.eps.numel(),
         sess.table.data_ptr(), row.data_ptr(), 0, st)
    for _ in range(3):
        check(fn(*a), "sched step")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        check(fn(*a), "sched step")
    e1.record()
    torch.cuda.synchronize(sess.dev)
    return e0.elapsed_time(e1) / reps


def kernel_classes(table, pk, sched_ms, numel):
    """per-kernel-class records of one eager, per-launch-timed denoise step: algorithmic work (the reference op count /
    the minimum bytes), executed work (what the kernels really do: sub-pixel upsample convs run 4/9 of the MACs, identity
    shortcuts ride along as GEMM panels), event time, fraction of the measured peak."""
    def sel(pred):
        return [(n, meta, m) for n, meta, m in table if pred(n)]
    out = {}
    conv = sel(lambda n: n == "conv")
    ms = sum(m for _, _, m in conv)
    fl = sum(meta["flops"] for _, meta, _ in conv)
    fx = sum(meta.get("flops_exec", meta["flops"]) for _, meta, _ in conv)
    out["conv"] = {"bound": "tensor", "launches": len(conv), "ms": ms, "algorithmic_flops": fl, "executed_flops": fx,
                   "achieved_tflops": fl / (ms * 1e-3) / 1e12, "frac": fl / (ms * 1e-3) / 1e12 / pk["tflops_sustained"],
                   "executed_tflops": fx / (ms * 1e-3) / 1e12,
                   "frac_executed": fx / (ms * 1e-3) / 1e12 / pk["tflops_sustained"]}
    # the cout = 64 and the upsample convs are the launches furthest from the roofline: listed on their own
    for key, pred in (("conv_cout64", lambda meta: meta["cout"] == 64 and meta["mode"] == 0),
                      ("conv_upsample", lambda meta: meta["mode"] == 2)):
        sub = [(meta, m) for _, meta, m in conv if pred(meta)]
        if sub:
            sms, sfl = sum(m for _, m in sub), sum(meta["flops"] for meta, _ in sub)
            sfx = sum(meta.get("flops_exec", meta["flops"]) for meta, _ in sub)
            out[key] = {"bound": "tensor", "launches": len(sub), "ms": sms, "achieved_tflops": sfl / (sms * 1e-3) / 1e12,
                        "frac": sfl / (sms * 1e-3) / 1e12 / pk["tflops_sustained"],
                        "frac_executed": sfx / (sms * 1e-3) / 1e12 / pk["tflops_sustained"]}
    for key, names in (("gn_apply", ("gn_apply",)), ("gn_stats", ("gn_stats",)), ("conv_in", ("conv_in",)),
                       ("conv_out", ("conv_out",))):
        rows = sel(lambda n: n in names)
        if rows:
            ms = sum(m for _, _, m in rows)
            by = sum(meta.get("bytes", 0) for _, meta, _ in rows)
            out[key] = {"bound": "hbm", "launches": len(rows), "ms": ms, "algorithmic_bytes": by,
                        "achieved_gbs": by / (ms * 1e-3) / 1e9, "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
    att = sel(lambda n: n == "attention")
    if att:
        ms = sum(m for _, _, m in att)
        fl = sum(meta["flops"] for _, meta, _ in att)
        ex = sum(meta.get("exps", 0) for _, meta, _ in att)
        out["attention"] = {"bound": "mufu (exp)", "launches": len(att), "ms": ms, "algorithmic_flops": fl, "exps": ex,
                            "achieved_tflops": fl / (ms * 1e-3) / 1e12, "gexps_per_s": ex / (ms * 1e-3) / 1e9}
    by = 16 * numel
    out["sched_step"] = {"bound": "hbm", "launches": 1, "ms": sched_ms, "algorithmic_bytes": by,
                         "achieved_gbs": by / (sched_ms * 1e-3) / 1e9,
                         "frac": by / (sched_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "note": "3 reads + 1 write of the fp32 sample; timed alone, so it partly runs out of the 126 MB L2"}
    other = sel(lambda n: n not in ("conv", "gn_apply", "gn_stats", "conv_in", "conv_out", "attention"))
    out["other"] = {"launches": len(other), "ms": sum(m for _, _, m in other),
                    "ops": sorted({n for n, _, _ in other})}
    return out


def pipeline_e2e(model, sched_cls, B, S, K, dev):
    """`DDPMPipeline.__call__` — the call the reference makes (training_pipeline.py:26-32, generation.py:14-20) — for K
    inference steps, host numpy out: with generator=None (generation.py) and with a CPU generator (evaluate())."""
    from drivescenegen_b200.hostapi import DDPMPipeline
    pipe = DDPMPipeline(unet=model, scheduler=sched_cls())
    pipe.set_progress_bar_config(disable=True)
    out = {}
    for key, mk in (("generator_none", lambda: None), ("generator_cpu", lambda: torch.Generator().manual_seed(1234))):
        pipe(batch_size=B, generator=mk(), num_inference_steps=K, output_type="np.array")   # graph capture + warm-up
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        img = pipe(batch_size=B, generator=mk(), num_inference_steps=K, output_type="np.array").images
        dt = time.perf_counter() - t0
        assert img.shape == (B, S, S, 3)
        nb = B * 3 * S * S * 4
        out[key] = {"value": B * K / dt, "unit": UNIT, "ms_per_step": 1000.0 * dt / K, "steps": K,
                    "h2d_bytes_per_step": nb if key == "generator_cpu" else 0, "d2h_bytes_total": nb,
                    "note": ("variance noise drawn by torch.randn with the caller's CPU generator every step (upstream RNG "
                             "rule) into pinned memory, copied on a side stream: host-RNG bound"
                             if key == "generator_cpu" else "noise drawn on the device; one D2H of the final images")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dsg", choices=["dsg", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--train-batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[2] training sub-record of the default line")
    ap.add_argument("--profile-out", default=None, help="write the per-launch table (JSON) here")
    ap.add_argument("--scheduler", default="ddpm", choices=["ddpm", "ddim"],
                    help="ddim = BASELINE configs[3] when combined with --size 512 --batch 8 (50-step DDIM, eta 0)")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample = BASELINE configs[1] (the headline metric) + a `train` sub-record (configs[2]); "
                         "train = configs[2] alone (fwd+bwd+AdamW)")
    args = ap.parse_args()
    if args.workload == "train":
        return run_train(args)
    if args.impl == "reference":
        return run_reference(args)

    from drivescenegen_b200 import _lib
    from drivescenegen_b200.hostapi import DDIMScheduler, DDPMScheduler, DenoiseSession, UNet2DModel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (the product has no CPU path)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=dev)
    assert _lib.load().dsg_device_ok() == 1, "sm_100 device required"
    W, K, B, S = max(3, args.warmup), max(1, args.steps), args.batch, args.size

    torch.manual_seed(0)  # identical random-init weights on every rank
    model = UNet2DModel(sample_size=(S, S), **REF_CFG).to(dev).eval()
    ddim = args.scheduler == "ddim"
    sched = DDIMScheduler() if ddim else DDPMScheduler()
    sched.set_timesteps(50 if ddim else 1000)
    shape = (B, 3, S, S)
    sess = DenoiseSession(model, sched, shape)
    gen = torch.Generator().manual_seed(1234 + rank)  # per-rank seed: independent replicas
    x0 = torch.randn(shape, generator=gen)
    noise_host = [torch.randn(shape, generator=gen).pin_memory() for _ in range(4)]
    noise_dev = [z.to(dev) for z in noise_host]
    out_host = torch.empty(shape).pin_memory()
    ts = [int(t) for t in sched.timesteps]

    def sched_ts(first, n):
        return [ts[(first + i) % len(ts)] for i in range(n)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ device-resident timing
    # one graph launch per step: the timestep schedule lives on the device, the sample ping-pongs between two buffers;
    # the variance noise of step i is noise_dev[i % 4] (a device-to-device copy inside the timed region)
    sess.load(x0.to(dev))
    sess.begin(sched_ts(0, W))
    for i in range(W):
        sess.advance(noise_dev[i % 4])
    chunks = [sched_ts(W + c, min(sess.max_steps, K - c)) for c in range(0, K, sess.max_steps)]
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sess.begin(chunks[0])
    e0.record()
    i = 0
    for ci, chunk in enumerate(chunks):
        if ci:
            sess.begin(chunk)
        for _ in chunk:
            sess.advance(noise_dev[i % 4])
            i += 1
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None

    # ------------------------------------------------------------------ end-to-end (host buffers in, host buffer out)
    # serial form: copy in -> step -> copy out -> host sync, every step
    sess.load(x0.to(dev))
    for i in range(3):
        sess.step_from_host(ts[i], noise_host[i % 4], out_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        sess.step_from_host(ts[(3 + i) % len(ts)], noise_host[i % 4], out_host)
    barrier()
    e2e_serial_s = time.perf_counter() - t0
    # pipelined form (the one the pipeline uses with host-side noise): the same bytes cross PCIe every step and
    # every step's result is read on the host, with the copies on their own streams beside the next step's compute
    out_hosts = [out_host, torch.empty(shape).pin_memory()]
    seen = []
    sess.load(x0.to(dev))
    sess.run_from_host(sched_ts(0, 3), noise_host, out_hosts)
    barrier()
    t0 = time.perf_counter()
    for chunk in [sched_ts(3 + c, min(sess.max_steps, K - c)) for c in range(0, K, sess.max_steps)]:
        sess.run_from_host(chunk, noise_host, out_hosts, on_result=lambda i, o: seen.append(float(o[0, 0, 0, 0])))
    barrier()
    e2e_s = time.perf_counter() - t0
    assert len(seen) == K

    # the reference-facing call: DDPMPipeline.__call__ (rank 0 at N = 1 only: it is host-side work per replica)
    pipe_e2e = None
    if world == 1 and not ddim:
        pipe_e2e = pipeline_e2e(model, DDPMScheduler, B, S, min(K, 1000), dev)

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t[0].item(), t[1].item() / 1000.0

    line = None
    if rank == 0:
        # -------------------------------------------------------------- roofline of the dominant kernel (rank 0)
        pk = peaks()
        prog = model.engine().program(B, S, S)
        eps = torch.empty(shape, device=dev)
        tf = torch.full((B,), 500.0, device=dev)
        prog.run_timed(sess.x, tf, eps)  # warm
        table = median_table([prog.run_timed(sess.x, tf, eps) for _ in range(5)])
        kernels = kernel_classes(table, pk, time_sched_step(sess), x0.numel())
        cv = kernels["conv"]
        total_ms = sum(m for _, _, m in table)
        fwd_flops = sum(meta.get("flops", 0) for _, meta, _ in table)
        fwd_flops_exec = sum(meta.get("flops_exec", meta.get("flops", 0)) for _, meta, _ in table)
        # DRAM traffic of the same launches: `ncu --set full` capture of this code (tools/conv_traffic.py writes the file
        # from the capture and records the commit it was taken at)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", f"conv_traffic_b{B}_{S}.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = {k: tj.get(k) for k in ("source", "commit", "launches") if k in tj}
        roofline = {"bound": "tensor",
                    "kernel": "igemm_halo_kernel / igemm_kernel (all conv3x3/1x1/linear launches of one step)",
                    "achieved": cv["achieved_tflops"], "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": cv["frac"], "peak_source": pk["source"] + " bf16 sustained (cuBLAS)",
                    "achieved_executed": cv["executed_tflops"], "frac_executed": cv["frac_executed"],
                    "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu --set full)",
                    "traffic_source": traffic_src,
                    "launches": cv["launches"], "avg_launch_ms": cv["ms"] / cv["launches"],
                    "algorithmic_flops_per_step": cv["algorithmic_flops"], "share_of_step": cv["ms"] / total_ms}
        whole = {"flops": fwd_flops, "flops_executed": fwd_flops_exec, "ms": ms / K,
                 "frac": fwd_flops / (ms / K * 1e-3) / 1e12 / pk["tflops_sustained"],
                 "frac_executed": fwd_flops_exec / (ms / K * 1e-3) / 1e12 / pk["tflops_sustained"],
                 "note": "whole denoise step (graph replay incl. GroupNorm, attention, scheduler) against the tensor peak"}
        breakdown = {"conv_ms": cv["ms"], "groupnorm_ms": kernels.get("gn_apply", {}).get("ms", 0.0)
                     + kernels.get("gn_stats", {}).get("ms", 0.0),
                     "attention_ms": kernels.get("attention", {}).get("ms", 0.0), "eager_step_ms": total_ms,
                     "unet_fwd_flops": fwd_flops}
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            with open(args.profile_out, "w") as f:
                json.dump({"batch": B, "size": S, "table": [{"op": n, **{k: v for k, v in meta.items()}, "ms": m}
                                                            for n, meta, m in table]}, f, indent=1)
        # -------------------------------------------------------------- CPU baseline (bounded sample, rank 0, N=1)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, dt = oracle_steps_per_s(2, S, 3, 1)
            cores = torch.get_num_threads()
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"batch 2 x 3 steps of {S}x{S}x3 after 1 warm-up (oracle, fp32, {cores} threads, {dt:.1f} s)"}
        n_bytes = x0.numel() * 4
        line = {"metric": METRIC if S == 256 else f"denoise-steps/sec ({S}x{S}x3 raster)",
                "value": world * B * K / (ms * 1e-3), "unit": UNIT,
                "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "fp16 operands, fp32 accumulate (schedulers fp32)", "data": "synthetic",
                "config": {"workload": f"{S}x{S}x3 BEV raster, full U-Net (56.6M params, random init), "
                                       f"{'50-step DDIM (eta 0)' if ddim else 'DDPM'} sampling, "
                                       f"batch {B} per GPU, one CUDA-graph launch per denoise step",
                           "batch_per_gpu": B, "parallelism": f"replicas x{world} (no collective on the sampling path)",
                           "l2": "per-step working set (several GB of activations) >> 126 MB L2, no flush needed"},
                "batch_steps_per_s": world * K / (ms * 1e-3), "unet_fwd_ms_eager_sum": total_ms,
                "clocks": clocks,
                "e2e": {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n_bytes,
                        "d2h_bytes_per_step": n_bytes, "ms_per_step": 1000.0 * e2e_s / K,
                        "how": "DenoiseSession.run_from_host: pinned host noise in, host result out every step, copies "
                               "on side streams", "serial_ms_per_step": 1000.0 * e2e_serial_s / K},
                "e2e_pipeline": pipe_e2e,
                "gpu_launches": K * sess.launches_per_step,
                "roofline": roofline, "whole_step": whole, "kernels": kernels, "breakdown": breakdown,
                "cpu_baseline": cpu}
    # ------------------------------------------------------------------ configs[2] in the same line (every rank runs it)
    if not args.no_train and S == 256 and not ddim:
        del sess
        torch.cuda.empty_cache()
        rec = train_record(args, dev, rank, world, max(5, min(K, 10)), 3,
                           world == 1 and not args.no_cpu_baseline)
        if rank == 0:
            line["train"] = rec
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
