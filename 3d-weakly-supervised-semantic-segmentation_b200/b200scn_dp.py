"""Scene-sharded data parallelism: one process per GPU, each rank builds its own rulebooks for its own
scenes; the only exchange is one fp32 gradient all-reduce (mean) per step over NCCL / NVLink, between
loss.backward() and optimizer.step() (train.py:80-81).  The reference has no distributed code at all; BatchNorm
statistics stay per rank, as independent scn replicas would behave.
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """The step's single exchange: every parameter gradient is packed into ONE flat fp32 buffer (one multi-tensor copy),
    all-reduced once -- sized for launch latency, not one collective per tensor -- and unpacked in place.
    (Gradients are NOT kept as views of the flat buffer: autograd would then add into them, one extra kernel per
    parameter and step.)"""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=self.params[0].device)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]

    def zero(self):
        for p in self.params:
            p.grad = None

    def allreduce_mean(self, group=None):
        world = dist.get_world_size(group)
        if world <= 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)
        for p, g in zip(self.params, grads):
            if p.grad is None:
                p.grad = g
        torch._foreach_copy_(grads, self.views)


def shard_scenes(n_scenes, rank, world):
    """Scenes r, r+world, ... go to rank r (SURVEY 8e partitioning)."""
    return list(range(rank, n_scenes, world))
