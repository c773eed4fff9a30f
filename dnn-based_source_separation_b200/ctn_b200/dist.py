"""Multi-GPU host logic for the separation path: one process per GPU, contiguous batch shards, no data-path
collective (every mixture is independent: gLN and PIT statistics are per sample -- SURVEY.md 8e).  The only
collectives are the scalar loss mean and the max-over-ranks step time.  Works with NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of a global batch; remainders go to the lowest ranks."""
    if global_batch < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_loss_mean(loss_b: torch.Tensor, global_batch: int) -> torch.Tensor:
    """Mean of per-sample losses over the GLOBAL batch (pit.py:41-42 semantics under sharding)."""
    s = loss_b.double().sum().reshape(1)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return (s / global_batch).float()[0]


def max_over_ranks(value: float, device=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value: float, device=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def allreduce_gradients(model, local_batch: int | None = None, global_batch: int | None = None) -> int:
    """Data-parallel gradient reduction for the batch-sharded training step (ONE all-reduce of ~20 MB per
    step; replaces the reference's nn.DataParallel gather/reduce, egs/wsj0-mix/conv-tasnet/local/train.py:95).

    Every rank holds d(mean loss over ITS shard)/dθ.  The gradient of the mean over the GLOBAL batch is
    Σ_r (local_batch_r / global_batch) · g_r, i.e. a plain average when shards are equal (the reference's batch mean,
    src/criterion/pit.py:41-42).  The native backward writes all gradients into one flat buffer
    (``model.last_flat_grad``); when the parameters' ``.grad`` still alias it, the collective runs in place on that single
    bucket, otherwise the gradients are flattened, reduced and copied back.  Returns the number of elements reduced."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    params = [p for p in model.parameters() if p.grad is not None]
    if not params:
        return 0
    scale = (float(local_batch) / float(global_batch)) if (local_batch is not None and global_batch) else 1.0 / world
    flat = getattr(model, "last_flat_grad", None)
    in_place = False
    if flat is not None and flat.device == params[0].grad.device:
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        in_place = all(lo <= p.grad.data_ptr() < hi and p.grad.is_contiguous() for p in params)
    if in_place:
        bucket = flat
    else:
        bucket = torch.cat([p.grad.reshape(-1) for p in params])
    if world > 1:
        bucket.mul_(scale)
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    if not in_place and world > 1:
        off = 0
        for p in params:
            n = p.grad.numel()
            p.grad.copy_(bucket[off:off + n].view_as(p.grad))
            off += n
    return bucket.numel()
