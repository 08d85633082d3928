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
                         src: int = 0, stats: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Rank `src` holds `sd`; every rank returns the same tensors on `device`.

    All tensors are packed into ONE flat byte buffer and sent with a single broadcast (NCCL over NVLink on the GPU
    box, gloo in the CPU tests).  Parameters the arenas keep as fp16 (Linear weights / biases: checkpoint.
    consumed_as_fp16) travel as fp16 -- the bits `tensor.half()` gives, which is all the kernels ever see -- the rest as
    fp32: 1.3 GB instead of 2.4 GB for the full model.  `specs` (name -> (shape, ...)) fixes the order.  `stats`
    (optional dict) receives the byte count and, on CUDA, the device time of the collective.
    """
    from .checkpoint import consumed_as_fp16
    names: List[str] = list(specs.keys())
    shapes = [tuple(specs[n][0]) for n in names]
    dtypes = [torch.float16 if consumed_as_fp16(n) else torch.float32 for n in names]
    offs, total = [], 0
    for shp, dt in zip(shapes, dtypes):
        offs.append(total)
        nbytes = int(torch.Size(shp).numel()) * (2 if dt == torch.float16 else 4)
        total += (nbytes + 255) // 256 * 256                       # kernels use 16-byte loads: keep every view aligned
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    flat = torch.empty(total, dtype=torch.uint8, device=device)

    def view(i):
        n = int(torch.Size(shapes[i]).numel())
        nbytes = n * (2 if dtypes[i] == torch.float16 else 4)
        return flat[offs[i]:offs[i] + nbytes].view(dtypes[i]).view(shapes[i])

    if rank == src:
        assert sd is not None
        for i, n in enumerate(names):
            view(i).copy_(sd[n].to(dtypes[i]))
    ms = None
    if world > 1:
        if device.type == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.broadcast(flat, src=src)
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1)
        else:
            dist.broadcast(flat, src=src)
    if stats is not None:
        stats.update(bytes=total, ms=ms, fp16_tensors=sum(d == torch.float16 for d in dtypes),
                     fp32_tensors=sum(d == torch.float32 for d in dtypes))
    return {n: view(i) for i, n in enumerate(names)}


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of the batch for this rank (accelerate's DataLoader sharding, main.py:146)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
