"""Multi-GPU plan for the generator forward path: shard the utterance batch, broadcast the weights once.

The path has no exchange step (utterances are independent, weights are read-only), so the only collective is
ONE broadcast of the packed folded-weight buffer at init (rank 0 loads the checkpoint; 12.7-18.6 MB fp32 over
NVLink/NVSwitch via NCCL) and there is no per-step collective.  One process per GPU (torchrun).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int):
    """Contiguous [lo, hi) utterance range of `rank`; the first batch % world ranks take one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment. Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def broadcast_weights(model, src: int = 0):
    """The single init-time collective: rank `src`'s packed folded weights -> every rank."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if getattr(model, "_wn", None):
            model.remove_weight_norm()          # fold on every rank so the packed buffer is what is sent
        dist.broadcast(model.packed_weights, src=src)
        model.packed_weights._version  # noqa: B018  (in-place broadcast bumps the version -> re-bind on next forward)
        model._bound_key = None
    return model


def max_over_ranks(value: float, device=None) -> float:
    """Device-side timing reduction: max over ranks of a per-rank elapsed time."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
