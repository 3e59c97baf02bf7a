"""Multi-GPU sharding of the decode path: one process per GPU, images partitioned contiguously, NO collective inside
the DDIM loop (every image is an independent chain — GroupNorm/FiLM are per sample).  After the loop:
one all_gather of the reconstructions and one all_reduce(SUM) of the fp64 metric sums — NCCL over NVLink on GPUs,
gloo in the CPU tests.  The reference has no distributed code at all; this is new capability."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous partition of range(n): the first n % world ranks hold one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the default process group if world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


def gather_shards(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all_gather of per-rank shards [n_r, ...] produced with shard_bounds -> [n_total, ...] on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = (n_total + world - 1) // world  # equal-size buffers (NCCL all_gather needs them)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(bufs[r][: hi - lo])
    assert parts[rank].shape[0] == local.shape[0]
    return torch.cat(parts, dim=0)


def reduce_sums(values, device) -> torch.Tensor:
    """all_reduce(SUM) of a small fp64 vector, e.g. [sum_psnr, n_finite_psnr, sum_ssim, n_finite_ssim]."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
