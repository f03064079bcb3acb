"""World-size-2 gloo test of the data-parallel contract of the `accelerate` shim:
gradients are averaged over ranks with ONE all-reduce on a flat buffer, dataloader batches are sharded round-robin,
the LR scheduler advances num_processes times per step, replicas stay bit-identical."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json, torch
    sys.path.insert(0, os.environ["DSG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DSG_ROOT"], "shims"))
    import torch.distributed as dist
    from accelerate import Accelerator
    calls = {"n": 0}
    orig = dist.all_reduce
    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    dist.all_reduce = counting
    acc = Accelerator(mixed_precision="no", gradient_accumulation_steps=1, cpu=True)
    rank, world = acc.process_index, acc.num_processes
    torch.manual_seed(100 + rank)            # deliberately different init per rank: prepare() must broadcast rank 0's
    model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    data = torch.arange(6 * 2 * 4, dtype=torch.float32).reshape(12, 4) / 10.0
    loader = torch.utils.data.DataLoader(data, batch_size=2, shuffle=False)
    model, opt, loader, sched = acc.prepare(model, opt, loader, sched)
    seen = []
    for batch in loader:
        seen.append(batch[:, 0].tolist())
        with acc.accumulate(model):
            loss = model(batch).pow(2).mean()
            acc.backward(loss)
            acc.clip_grad_norm_(model.parameters(), 1.0)
            opt.step(); sched.step(); opt.zero_grad()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    # one file per rank: two ranks sharing one stdout pipe can interleave inside a line
    with open(os.path.join(os.environ["DSG_OUT"], f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "seen": seen, "w": flat.tolist(), "lr": sched.get_last_lr()[0],
                   "allreduce_calls": calls["n"], "len": len(loader)}, f)
    dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_gradient_averaging(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DSG_ROOT=ROOT, DSG_OUT=str(tmp_path), CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    # round-robin sharding: rank 0 sees batches 0,2,4 and rank 1 sees 1,3,5 (first column identifies the rows)
    assert res[0]["len"] == 3 and len(res[0]["seen"]) == 3 and len(res[1]["seen"]) == 3
    assert res[0]["seen"][0][0] == 0.0 and abs(res[1]["seen"][0][0] - 0.8) < 1e-6
    # replicas identical after training (same averaged gradients applied to the same broadcast init)
    assert res[0]["w"] == res[1]["w"]
    # one all-reduce per step (3 steps); barrier() may add its own, so count >= 3 and <= 3 + few
    assert 3 <= res[0]["allreduce_calls"] <= 5
    # upstream quirk: scheduler advanced num_processes (=2) times per step -> after 3 steps lr = 1/(1+6)
    assert abs(res[0]["lr"] - 0.1 / 7) < 1e-9

    # reference result computed single-process: average of the two per-rank gradients == gradient of the mean loss
    import torch
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    data = torch.arange(6 * 2 * 4, dtype=torch.float32).reshape(12, 4) / 10.0
    for step in range(3):
        b0, b1 = data[4 * step:4 * step + 2], data[4 * step + 2:4 * step + 4]
        loss = 0.5 * (model(b0).pow(2).mean() + model(b1).pow(2).mean())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step(); sched.step(); sched.step(); opt.zero_grad()
    ref = torch.cat([p.detach().flatten() for p in model.parameters()])
    got = torch.tensor(res[0]["w"])
    assert torch.allclose(got, ref, atol=1e-6), (got - ref).abs().max()


SHARD_WORKER = textwrap.dedent("""
    import os, sys, json, torch
    sys.path.insert(0, os.environ["DSG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DSG_ROOT"], "shims"))
    import torch.distributed as dist
    from accelerate import Accelerator

    class Counting(torch.utils.data.Dataset):
        def __init__(self, n): self.n, self.loaded = n, []
        def __len__(self): return self.n
        def __getitem__(self, i):
            self.loaded.append(int(i))
            return torch.tensor([float(i)])

    acc = Accelerator(cpu=True)
    rank = acc.process_index
    ds = Counting(11)                       # 11 samples, batch 2 -> 6 batches (last one short), 2 ranks -> 3 steps each
    loader = acc.prepare(torch.utils.data.DataLoader(ds, batch_size=2, shuffle=True))
    epochs = []
    for epoch in range(3):
        ds.loaded.clear()
        seen = [b.flatten().tolist() for b in loader]
        epochs.append({"seen": seen, "loaded": sorted(ds.loaded)})
        if rank == 0:
            torch.manual_seed(14555)        # the reference's evaluate() reseeds the global RNG on rank 0 only
        else:
            torch.rand(7)                   # ... and other ranks drift differently
    with open(os.path.join(os.environ["DSG_OUT"], f"rank{rank}.json"), "w") as f:
        json.dump({"rank": rank, "epochs": epochs, "len": len(loader)}, f)
    dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_loader_shards_at_the_sampler_and_keeps_shuffles_in_sync(tmp_path):
    """each rank loads ONLY its own samples (sharding at the batch-sampler level), the shuffle generator is
    re-synchronised from rank 0 every epoch (ranks' global RNG states diverge between epochs), every rank runs the same
    number of full-size batches, and the tail of the epoch is padded from the epoch's first samples."""
    import json
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER)
    env = dict(os.environ, DSG_ROOT=ROOT, DSG_OUT=str(tmp_path), CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29633", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    assert res[0]["len"] == res[1]["len"] == 3
    orders = []
    for e in range(3):
        a, b = res[0]["epochs"][e], res[1]["epochs"][e]
        assert [len(x) for x in a["seen"]] == [2, 2, 2] and [len(x) for x in b["seen"]] == [2, 2, 2]
        # a rank decodes exactly the samples it trains on — not the whole dataset
        assert sorted(int(v) for s in a["seen"] for v in s) == a["loaded"]
        assert sorted(int(v) for s in b["seen"] for v in s) == b["loaded"]
        inter = [v for pair in zip(a["seen"], b["seen"]) for s in pair for v in s]   # the global batch order
        assert sorted(set(inter)) == [float(i) for i in range(11)]                   # every sample seen each epoch
        assert len(inter) == 12 and inter[11] == inter[0]                            # short batch topped up from the head
        orders.append(inter)
    assert orders[0] != orders[1] or orders[1] != orders[2]                          # it does shuffle


def test_batch_sampler_shard_padding_rules():
    from drivescenegen_b200.hostapi.accelerator import BatchSamplerShard
    from torch.utils.data import BatchSampler, SequentialSampler

    def shards(n, bs, world, drop_last=False):
        base = BatchSampler(SequentialSampler(range(n)), bs, drop_last)
        out = [list(BatchSamplerShard(base, r, world)) for r in range(world)]
        assert all(len(o) == len(BatchSamplerShard(base, r, world)) for r, o in enumerate(out))
        return out

    assert shards(8, 2, 2) == [[[0, 1], [4, 5]], [[2, 3], [6, 7]]]
    # 3 batches for 2 ranks: the missing 4th batch comes from the head of the epoch
    assert shards(6, 2, 2) == [[[0, 1], [4, 5]], [[2, 3], [0, 1]]]
    # short last batch is topped up, then the round is completed
    assert shards(5, 2, 2) == [[[0, 1], [4, 0]], [[2, 3], [1, 2]]]
    assert shards(7, 2, 2) == [[[0, 1], [4, 5]], [[2, 3], [6, 0]]]
    assert shards(7, 2, 2, drop_last=True) == [[[0, 1]], [[2, 3]]]
    assert shards(3, 2, 4) == [[[0, 1]], [[2, 0]], [[1, 2]], [[0, 1]]]
