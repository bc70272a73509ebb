"""CPU suite for the N > 1 host logic (world_size 2, gloo): scene sharding and the single flat gradient
all-reduce of the training step keep replicas identical.  The model itself needs a GPU (no CPU path), so
a small stand-in module exercises the same Trainer code."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from s4g_release_b200.train import GradBucket, Trainer, shard_range


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.bn = torch.nn.BatchNorm1d(5)
        self.b = torch.nn.Linear(5, 3)

    def forward(self, batch):
        return {"y": self.b(torch.relu(self.bn(self.a(batch["x"]))))}


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    model = _Toy()
    trainer = Trainer(model, lambda p, l: {"mse": ((p["y"] - l["t"]) ** 2).mean()}, lr=1e-2)
    g = torch.Generator().manual_seed(7)
    x_all, t_all = torch.randn(16, 6, generator=g), torch.randn(16, 3, generator=g)
    lo, hi = shard_range(16, rank, world)
    losses = []
    for _ in range(5):
        losses.append(trainer.step({"x": x_all[lo:hi]}, {"t": t_all[lo:hi]})["mse"].item())
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out["sync"] = all(torch.equal(gathered[0], g_) for g_ in gathered)
        out["losses"] = losses
    dist.destroy_process_group()


def test_two_rank_training_step_keeps_replicas_in_sync():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert out["sync"], "replicas diverged after all-reduced steps"
        assert out["losses"][-1] < out["losses"][0]


def test_grad_bucket_aliases_parameter_grads():
    model = _Toy()
    bucket = GradBucket(model.parameters())
    loss = model({"x": torch.randn(4, 6)})["y"].sum()
    loss.backward()
    assert bucket.flat.abs().sum() > 0
    n = sum(p.numel() for p in model.parameters())
    assert bucket.flat.numel() == n
    bucket.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in model.parameters())
