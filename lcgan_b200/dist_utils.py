"""Host side of the single-node data-parallel path (reference main.py:105-113, worker.py:35,88-96): one
process per GPU, batch sharded across ranks, gradients averaged over NCCL.  LC-GAN's only exchange step
is the gradient all-reduce.  Two drivers:
  * torch DDP, as the reference wraps G and D (`wrap_ddp`): our backward kernels run on the current
    stream, so DDP's bucket-ready events order correctly and its buckets overlap with them;
  * `GradExchange`: the same bucketed, backward-overlapped all-reduce without the DDP wrapper, built from
    post-accumulate-grad hooks and a side stream, so that it can be captured into a CUDA graph together
    with the kernels (DDP's reducer cannot)."""
import torch
import torch.distributed as dist


def local_batch(global_batch: int, world: int) -> int:
    """worker.py:35: local_batch = batch_size // gpus (the remainder is dropped, like the reference)."""
    if world < 1 or global_batch < world:
        raise ValueError(f"global batch {global_batch} cannot be sharded over {world} ranks")
    return global_batch // world


def rank_seed(rank: int, base: int = 1000) -> int:
    """Distinct synthetic-data stream per rank."""
    return base + rank


def max_over_ranks(value: float, device=None) -> float:
    """Device time of a multi-GPU step is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def wrap_ddp(module, device_ids=None):
    """DistributedDataParallel exactly as the reference wraps G and D (worker.py:88-96)."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    return DDP(module, device_ids=device_ids, broadcast_buffers=False, find_unused_parameters=True)


class _BucketPlan:
    """Gradient buckets of one step variant, in the order backward produces the gradients."""

    def __init__(self):
        self.order, self.recording = [], True
        self.slot, self.flats, self.members, self.pending = {}, [], [], []

    def build(self, arena, cap_bytes):
        """Carve the recorded parameters into buckets of <= cap_bytes inside `arena` (one flat f32 buffer
        shared by all variants of a network: only one variant runs at a time)."""
        self.recording = False
        off, start, members = 0, 0, []

        def close():
            if members:
                self.flats.append(arena[start:off])
                self.members.append(list(members))
        for p in self.order:
            n = p.numel()
            if members and (off - start + n) * 4 > cap_bytes:
                close()
                start, members = off, []
            self.slot[p] = (len(self.flats), off - start)
            members.append(p)
            off += n
        close()
        self.pending = [len(m) for m in self.members]


class GradExchange:
    """Bucketed gradient all-reduce (mean) overlapped with backward, for a fixed set of step variants.

    `backward(loss, variant, net)` runs loss.backward() with every parameter's post-accumulate-grad hook
    copying the fresh gradient into its slot of a flat bucket; the moment a bucket is complete its
    all-reduce is issued on a side stream (CUDA) while backward continues.  Afterwards each p.grad IS its
    averaged bucket slice (no copy back).  The first call for a variant records the order in which
    gradients appear (it is static per variant) and exchanges after backward; parameters that receive no
    gradient stay at grad None, as under DDP(find_unused_parameters=True).  Everything is issued on
    streams, never synchronises the host, and is CUDA-graph capturable."""

    def __init__(self, nets, device, world, bucket_bytes=48 << 20):
        self.world, self.bucket_bytes = world, bucket_bytes
        self.cuda = torch.device(device).type == "cuda"
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None
        # NCCL averages inside the collective; gloo has no AVG
        self.avg = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else None
        self.arena, self.plans, self.active = {}, {}, None
        for name, net in nets.items():
            params = list(net.parameters())
            self.arena[name] = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=device)
            for p in params:
                p.register_post_accumulate_grad_hook(self._hook)

    def _all_reduce(self, t):
        if self.avg is not None:
            dist.all_reduce(t, op=self.avg)
        else:
            dist.all_reduce(t)
            t.div_(self.world)

    def _hook(self, p):
        plan = self.active
        if plan is None:
            return
        if plan.recording:
            plan.order.append(p)
            return
        if p not in plan.slot:                        # not seen when the variant was recorded: exchange it alone
            return self._all_reduce(p.grad)
        b, off = plan.slot[p]
        view = plan.flats[b][off:off + p.numel()].view_as(p)
        view.copy_(p.grad)
        p.grad = view                                 # the optimizer reads the averaged gradient in place
        plan.pending[b] -= 1
        if plan.pending[b] == 0:
            self._launch(plan, b)

    def _launch(self, plan, b):
        if not self.cuda:
            return self._all_reduce(plan.flats[b])
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self._all_reduce(plan.flats[b])

    def backward(self, loss, variant, net):
        plan = self.plans.get(variant)
        if plan is None:
            plan = self.plans[variant] = _BucketPlan()
        self.active = plan
        try:
            loss.backward()
        finally:
            self.active = None
        if plan.recording:
            for p in plan.order:
                self._all_reduce(p.grad)
            plan.build(self.arena[net], self.bucket_bytes)
            return
        for b, left in enumerate(plan.pending):      # a bucket some parameter never reported to
            if 0 < left < len(plan.members[b]):
                self._launch(plan, b)
        plan.pending = [len(m) for m in plan.members]
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm)
