"""Multi-GPU plumbing: one process per GPU, coalition rows sharded, one all-gather of the outputs.

Coalitions are independent evaluations of one clip, so the data path has no
collective; the single exchange is an all-gather of the per-coalition output rows before the
regression.  Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as td


def init_from_env(device_type: str = "cuda"):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  No-op for one process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or td.is_initialized():
        return rank_world()
    backend = "nccl" if device_type == "cuda" else "gloo"
    if device_type == "cuda":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    td.init_process_group(backend=backend)
    return rank_world()


def rank_world():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_range(num_rows: int, rank: int, world: int):
    """Contiguous row block of ``rank``: ceil(K / G) rows per rank, the last ranks may be short or empty."""
    per = (num_rows + world - 1) // world
    lo = min(rank * per, num_rows)
    hi = min(lo + per, num_rows)
    return lo, hi


def all_gather_rows(y_local: torch.Tensor, num_rows: int, rank: int, world: int) -> torch.Tensor:
    """Gather the row blocks of every rank into the full [num_rows, D] matrix (identical on all ranks)."""
    if world == 1:
        return y_local
    per = (num_rows + world - 1) // world
    D = y_local.shape[1]
    padded = torch.zeros((per, D), dtype=y_local.dtype, device=y_local.device)
    padded[: y_local.shape[0]] = y_local
    out = torch.empty((world * per, D), dtype=y_local.dtype, device=y_local.device)
    td.all_gather_into_tensor(out, padded)
    return out[:num_rows]


def barrier():
    if td.is_available() and td.is_initialized():
        td.barrier()
