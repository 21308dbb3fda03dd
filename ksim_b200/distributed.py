"""One process per GPU: environment sharding and the (few) cross-rank reductions of the rollout path.

Environments are independent, so the data path has NO collective: rank r owns the contiguous slice
``shard_range(total_envs, r, world)`` and derives the same per-env keys it would get in a single-process run
(``make_env_init(env_offset=...)``), which makes results invariant to the GPU count.  The only exchanges are
scalar reductions for reporting (max time over ranks, summed env-steps, optional global advantage statistics).
Backend: ``nccl`` on GPUs, ``gloo`` in the CPU tests.
"""

from __future__ import annotations

__all__ = ["shard_range", "init_from_env", "shutdown", "max_over_ranks", "sum_over_ranks", "global_mean_std", "allreduce_mean_"]

import datetime
import os

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of ``total`` items for ``rank``."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None, timeout_s: float | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group when world > 1.  ``timeout_s``: the
    collective watchdog's timeout (a benchmark wants a mismatched collective to fail in minutes, not after the default 10)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {} if timeout_s is None else {"timeout": datetime.timedelta(seconds=float(timeout_s))}
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local), **kw)
        else:
            dist.init_process_group(backend, **kw)
    return rank, local, world


def _reduce(value: float, op: dist.ReduceOp, device: torch.device | str | None) -> float:
    if device is None:  # NCCL reduces device tensors only: follow the backend, not a fixed default
        nccl = dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
        device = torch.device("cuda", torch.cuda.current_device()) if nccl else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value: float, device: torch.device | str | None = None) -> float:
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value: float, device: torch.device | str | None = None) -> float:
    return _reduce(value, dist.ReduceOp.SUM, device)


def global_mean_std(x: torch.Tensor) -> tuple[float, float]:
    """Mean / std of a tensor sharded over ranks (for job-wide advantage normalisation): one 3-float all-reduce."""
    acc = torch.stack([x.sum(dtype=torch.float64), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc)
    mean = acc[0] / acc[2]
    var = torch.clamp(acc[1] / acc[2] - mean * mean, min=0.0)
    return float(mean), float(var.sqrt())


def allreduce_mean_(tensors: list[torch.Tensor], bucket_bytes: int = 64 << 20) -> None:
    """In-place mean over ranks of a list of tensors -- the one collective of the data-parallel PPO step
    (ksim/task/ppo.py:530-533 computes the gradient of a rank's minibatch; with env-sharded rollouts the ranks'
    gradients are averaged before the optimizer update).

    Tensors are packed into flat buckets (one ``all_reduce`` each: NVSwitch cost is launch-latency bound, not
    per-link bound, so few large buckets) and unpacked in place.  No-op in a single-process run.
    """
    if not tensors or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    bucket: list[torch.Tensor] = []
    size = 0

    def flush() -> None:
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat)
        flat.div_(world)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off : off + n].view_as(t))
            off += n
        bucket, size = [], 0

    for t in tensors:
        if bucket and (t.dtype != bucket[0].dtype or size + t.numel() * t.element_size() > bucket_bytes):
            flush()
        bucket.append(t)
        size += t.numel() * t.element_size()
    flush()


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
