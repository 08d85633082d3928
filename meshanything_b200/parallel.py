"""One process per GPU; shapes of a batch are independent: the only collective is the
weight broadcast at init (the reference gets the same from DDP's constructor, main.py:146)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank); initialises torch.distributed when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], specs: Dict[str, tuple], device: torch.device,
                         src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` holds `sd`; every rank returns the same tensors on `device`.

    All tensors are packed into ONE flat fp32 buffer and sent with a single broadcast (NCCL over
    NVLink on the GPU box, gloo in the CPU tests).  `specs` (name -> (shape, ...)) fixes the order.
    """
    names: List[str] = list(specs.keys())
    shapes = [tuple(specs[n][0]) for n in names]
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    total = sum(sizes)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if rank == src:
        assert sd is not None
        off = 0
        for n, sz in zip(names, sizes):
            flat[off:off + sz].copy_(sd[n].reshape(-1).to(torch.float32))
            off += sz
    if world > 1:
        dist.broadcast(flat, src=src)
    out, off = {}, 0
    for n, shp, sz in zip(names, shapes, sizes):
        out[n] = flat[off:off + sz].view(shp)
        off += sz
    return out


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of the batch for this rank (accelerate's DataLoader sharding, main.py:146)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
