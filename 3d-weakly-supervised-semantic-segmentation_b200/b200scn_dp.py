"""Scene-sharded data parallelism: one process per GPU, each rank builds its own rulebooks for its own
scenes; the only exchange is the fp32 gradient all-reduce (mean) of every step over NCCL, between loss.backward() and
optimizer.step() (train.py:80-81).  The reference has no distributed code at all; BatchNorm statistics stay per rank,
as independent scn replicas would behave.
"""
import torch
import torch.distributed as dist


class FlatGrads:
    """The step's single exchange.  Every parameter gradient has a slot in ONE flat fp32 buffer, cut into a few buckets
    (reverse parameter order = the order backward produces them).  A post-accumulate hook per parameter counts a bucket
    down; the moment a bucket is complete its gradients are packed with one multi-tensor copy and its all-reduce starts
    asynchronously, so the transfer runs under the rest of the backward pass -- measured on 2 B200s of one box the
    un-overlapped 120 MB all-reduce cost 14 ms of a 58 ms step.  `allreduce_mean()` (after backward) reduces whatever is
    still pending, waits, scales and unpacks in place.
    (During backward gradients are NOT views of the flat buffer -- autograd would then add into them, one extra kernel per
    parameter and step; AFTER the all-reduce every .grad is re-pointed at its slice, so there is no unpack copy.)"""

    def __init__(self, params, bucket_bytes=16 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=self.params[0].device)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]
        self.group = group
        # buckets: contiguous parameter ranges, launched in reverse parameter order (= the order backward produces them).
        # The bucket that holds parameter 0 goes out LAST, when backward is over, and its all-reduce is fully exposed: the
        # boundaries are therefore cut from the FRONT with growing sizes (1/16, 1/4, then whole buckets), so that the last
        # collectives are small (cfg3: 1.3 MB and 4.8 MB instead of one 8.2 MB tail bucket; at N=8 the step waited 1.2 ms for it).
        ends = [0]
        for sz in self.sizes:
            ends.append(ends[-1] + sz)
        cuts, lo, acc, k = [], 0, 0, 0
        limits = [bucket_bytes // 16, bucket_bytes // 4]
        for i in range(len(self.params)):
            acc += self.sizes[i] * 4
            lim = limits[k] if k < len(limits) else bucket_bytes
            if acc >= lim or i == len(self.params) - 1:
                cuts.append((lo, i + 1, ends[lo], ends[i + 1]))
                lo, acc, k = i + 1, 0, k + 1
        self.buckets = cuts[::-1]      # (first_param, last_param_exclusive, flat_begin, flat_end), launch order
        self.bucket_of = {}
        for b, (lo, hi_, _, _) in enumerate(self.buckets):
            for i in range(lo, hi_):
                self.bucket_of[i] = b
        self._pending = None       # per bucket: parameters still missing this step
        self._works = []
        self._ready, self._next, self._ev = [], 0, None
        self._hooks = []
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then one division
        self.op = dist.ReduceOp.SUM
        if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
            self.op = dist.ReduceOp.AVG
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    # ------------------------------------------------------------------ per step
    def zero(self):
        """Start of a step (before forward).  One backward per step: a second backward without zero() is refused."""
        for p in self.params:
            p.grad = None
        self.zero_state_only()

    def zero_state_only(self):
        self._pending = [hi - lo for lo, hi, _, _ in self.buckets]
        self._works = []
        self._ready = [False] * len(self.buckets)
        self._next = 0             # buckets are launched strictly in index order on every rank (matching collectives)
        self._ev = None

    def _make_hook(self, i):
        def hook(_p):
            if self._pending is None:
                return
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] < 0:
                raise RuntimeError("FlatGrads: a parameter received a second gradient in one step; call zero() before "
                                   "every backward (one backward per step is supported)")
            if self._pending[b] == 0:
                self._ready[b] = True
                self._launch_ready()
        return hook

    def _launch_ready(self):
        # a bucket only goes out once every bucket before it has gone out: a rank whose buckets fill in another order
        # (data-dependent branches, parameters without a gradient) still issues the same collectives in the same order
        while self._next < len(self.buckets) and self._ready[self._next]:
            self._launch(self._next)
            self._next += 1

    def _launch(self, b):
        lo, hi, f0, f1 = self.buckets[b]
        side = self._deferred_stream()
        if side is not None:
            # deferred weight gradients (sparseconvnet.ops.set_deferred_dw): some gradients of the bucket are still being
            # computed on the second stream, the others on the main stream -- pack and launch from the second stream once it
            # has caught up with the main stream's position, so that the main stream never waits for a weight gradient
            side.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(side):
                self._pack_and_reduce(lo, hi, f0, f1)
        else:
            self._pack_and_reduce(lo, hi, f0, f1)

    def _deferred_stream(self):
        if not self.flat.is_cuda:
            return None
        try:
            from sparseconvnet import ops
        except Exception:
            return None
        return ops._side_stream(self.flat.device) if ops._dw_defer[0] else None

    def _pack_and_reduce(self, lo, hi, f0, f1):
        grads = [self.params[i].grad if self.params[i].grad is not None else torch.zeros_like(self.params[i])
                 for i in range(lo, hi)]
        torch._foreach_copy_(self.views[lo:hi], grads)
        # AVG: the 1/world scale happens inside the collective (no separate pass over the flat buffer)
        self._works.append(dist.all_reduce(self.flat[f0:f1], op=self.op, group=self.group, async_op=True))

    def allreduce_mean(self, group=None):
        """After backward: send what is still pending (in order), wait, and point every .grad at its slice of the reduced
        flat buffer (no unpack copy: the optimiser reads the flat buffer through the views).  Returns nothing; the time the
        compute stream spent waiting for the collectives is available from `exposed_ms()` after a synchronize."""
        world = dist.get_world_size(self.group if group is None else group)
        if world <= 1:
            return
        if self._pending is None:          # zero() was not called this step: everything is still to do
            self.zero_state_only()
        for b in range(self._next, len(self.buckets)):   # parameters without a gradient this step, or hooks not registered
            self._launch(b)
        self._next = len(self.buckets)
        if self.flat.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for w in self._works:
            w.wait()
        if self.flat.is_cuda:
            e1.record()
            self._ev = (e0, e1)
        if self.op != dist.ReduceOp.AVG:
            self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._pending = None

    def exposed_ms(self):
        """Milliseconds the compute stream waited for the step's collectives (CUDA events around the waits); call after
        torch.cuda.synchronize().  Everything else of the all-reduce ran under backward."""
        if self._ev is None:
            return 0.0
        return self._ev[0].elapsed_time(self._ev[1])


def shard_scenes(n_scenes, rank, world):
    """Scenes r, r+world, ... go to rank r (SURVEY 8e partitioning)."""
    return list(range(rank, n_scenes, world))
