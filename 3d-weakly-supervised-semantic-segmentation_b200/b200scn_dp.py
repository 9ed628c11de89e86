"""Scene-sharded data parallelism: one process per GPU, each rank builds its own rulebooks for its own
scenes; the only exchange is one fp32 gradient all-reduce (mean) per step over NCCL / NVLink, between
loss.backward() and optimizer.step() (train.py:80-81).  The reference has no distributed code at all; BatchNorm
statistics stay per rank, as independent scn replicas would behave.
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """All parameter gradients live in ONE flat fp32 buffer (each p.grad is a view), so the step's exchange is a
    single all-reduce sized for launch latency, not one collective per tensor."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None):
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)


def shard_scenes(n_scenes, rank, world):
    """Scenes r, r+world, ... go to rank r (SURVEY 8e partitioning)."""
    return list(range(rank, n_scenes, world))
