"""Synthetic inputs of SURVEY.md 8d: shape i is a seeded (4096,6) point cloud with unit normals,
normalised exactly as the reference's Dataset.__getitem__ does (/root/reference/main.py:45-58)."""
from __future__ import annotations

import numpy as np
import torch


def normalize_pc_normal(pc_normal: np.ndarray) -> np.ndarray:
    """main.py:47-55: centre on the bbox, scale to max|x| = 0.9995, check unit normals, cast fp16."""
    pc_coor = pc_normal[:, :3]
    normals = pc_normal[:, 3:]
    bounds = np.array([pc_coor.min(axis=0), pc_coor.max(axis=0)])
    pc_coor = pc_coor - (bounds[0] + bounds[1])[None, :] / 2
    pc_coor = pc_coor / np.abs(pc_coor).max() * 0.9995
    assert (np.linalg.norm(normals, axis=-1) > 0.99).all(), "normals should be unit vectors, something wrong"
    return np.concatenate([pc_coor, normals], axis=-1, dtype=np.float16)


def synthetic_pc_normal(batch: int, first: int = 0, n_points: int = 4096) -> torch.Tensor:
    """fp16 [batch, n_points, 6]; shape i uses torch.Generator().manual_seed(1000 + first + i)."""
    out = []
    for i in range(batch):
        g = torch.Generator().manual_seed(1000 + first + i)
        xyz = torch.rand(n_points, 3, generator=g, dtype=torch.float64) * 2 - 1
        nrm = torch.randn(n_points, 3, generator=g, dtype=torch.float64)
        nrm = nrm / nrm.norm(dim=-1, keepdim=True)
        out.append(torch.from_numpy(normalize_pc_normal(torch.cat([xyz, nrm], dim=-1).numpy())))
    return torch.stack(out)
