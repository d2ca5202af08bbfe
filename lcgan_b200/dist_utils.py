"""Host-side helpers for the single-node data-parallel path (reference main.py:105-113,
worker.py:35,88-96): one process per GPU, batch sharded across ranks, gradients averaged by torch
DDP over NCCL.  No data-path collective of our own: LC-GAN's only exchange step is the gradient
all-reduce, which DDP buckets and overlaps with our backward kernels (they run on the current
stream, so DDP's bucket-ready events order correctly)."""
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
