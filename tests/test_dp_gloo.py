"""Scene-sharded data parallelism on CPU: world_size 2 over gloo.  The flat-gradient all-reduce must give
every rank the mean of the per-rank gradients, and shards must partition the scenes."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200")


def _worker(rank, world, port, out):
    sys.path.insert(0, PKG)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200scn_dp import FlatGrads, shard_scenes
    torch.manual_seed(0)                       # identical initial weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    # 8-byte buckets: every parameter is its own bucket, reduced asynchronously from its gradient hook during backward
    flat = FlatGrads(net.parameters(), bucket_bytes=8)
    assert len(flat.buckets) == 4 and flat.buckets[0][1] == 4 and flat.buckets[-1][0] == 0
    scenes = shard_scenes(5, rank, world)      # rank 0: 0,2,4   rank 1: 1,3
    torch.manual_seed(100 + rank)
    x = torch.randn(4 * len(scenes), 6)
    flat.zero()
    local_parts = {}
    for i, p in enumerate(net.parameters()):   # capture the local gradient before the in-flight reduction touches it
        p.register_hook(lambda g, i=i: local_parts.__setitem__(i, g.detach().clone().reshape(-1)))
    net(x).pow(2).mean().backward()
    assert flat._next == len(flat.buckets)     # every bucket's all-reduce was started by a hook, inside backward
    local = torch.cat([local_parts[i] for i in range(4)])
    flat.allreduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    now = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    ok = torch.allclose(now, sum(gathered) / world, atol=1e-7)
    same = all(torch.equal(a, b) for a, b in zip(gathered, gathered)) and not torch.equal(gathered[0], gathered[1])
    # ranks whose buckets fill in different orders (a branch only rank 0 runs) still issue the same collectives in the same
    # order: no hang, and the missing gradient counts as zero
    torch.manual_seed(0)
    a, b, c = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)
    flat2 = FlatGrads(list(a.parameters()) + list(b.parameters()) + list(c.parameters()), bucket_bytes=8)
    flat2.zero()
    torch.manual_seed(7 + rank)
    h = a(torch.randn(3, 4))
    if rank == 0:
        h = h + b(h)
    c(h).sum().backward()
    gb_local = b.weight.grad.clone() if rank == 0 else torch.zeros_like(b.weight)
    flat2.allreduce_mean()
    gb = [torch.zeros_like(gb_local) for _ in range(world)]
    dist.all_gather(gb, gb_local)
    branch_ok = torch.allclose(b.weight.grad, sum(gb) / world, atol=1e-7)
    # a second backward in the same step is refused instead of silently reusing stale buckets
    flat2.zero()
    c(a(torch.randn(3, 4))).sum().backward()
    try:
        c(a(torch.randn(3, 4))).sum().backward()
        refused = False
    except RuntimeError:
        refused = True
    flat2.allreduce_mean()
    out[rank] = (ok and branch_ok and refused, same, scenes)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] and out[1][1]
    assert sorted(out[0][2] + out[1][2]) == [0, 1, 2, 3, 4]


def _worker_net(rank, world, port, out):
    """A real encoder (the CPU oracle's SparseConvUNet, one scene per rank): after the exchange every rank holds the mean of
    the per-shard single-process gradients, bit-identical across ranks."""
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from b200scn_dp import FlatGrads, shard_scenes
    from b200scn_synth import build_encoder, make_batch
    from oracle import scn_oracle as ref
    torch.manual_seed(0)                       # identical initial weights on every rank
    net = build_encoder(ref, "SparseConvUNet", 4, 1, True)
    flat = FlatGrads(net.parameters(), bucket_bytes=4096)
    scenes = shard_scenes(world, rank, world)  # one scene each
    coords, feats, _ = make_batch(scenes, 8, n_points=1500)
    flat.zero()
    net([coords, feats]).pow(2).mean().backward()
    local = torch.cat([p.grad.reshape(-1).clone() for p in net.parameters()])   # this shard's single-process gradient
    flat.allreduce_mean()
    now = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = sum(gathered) / world
    rel = float((now - mean).norm() / mean.norm())
    everyone = [torch.zeros_like(now) for _ in range(world)]
    dist.all_gather(everyone, now)
    out[rank] = (rel, all(torch.equal(everyone[0], e) for e in everyone), len(flat.buckets))
    dist.destroy_process_group()


def test_encoder_gradients_are_the_mean_of_the_shards():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_net, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        rel, identical, nb = out[r]
        assert rel < 1e-6, rel          # = the mean of the per-shard gradients (bar: 1e-5)
        assert identical                 # bit-identical on every rank
        assert nb > 3                    # several buckets, launched from hooks during backward
