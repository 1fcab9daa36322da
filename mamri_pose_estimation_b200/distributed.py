"""Scan sharding and the single exchange of the path: scans are independent (`MamriLogic.process`
handles exactly one inputVolume, Mamri/Mamri.py:850-858), so a batch shards across ranks with no
data-path collective; the per-scan marker tables are all-gathered once per batch (NCCL over NVLink on
GPUs; the same code runs on gloo/CPU tensors in the tests)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

TABLE_SLOTS = 32        # marker slots per scan
TABLE_FIELDS = 8        # label, count, volume_mm3, ras x, ras y, ras z, n_labels, body_label


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin: item i -> rank i % world."""
    return list(range(rank, n_items, world))


def pack_table(results: Sequence) -> np.ndarray:
    """Fixed-size table [S, TABLE_SLOTS, TABLE_FIELDS] float64 from DetectionResult-like objects
    (attributes: markers[label,count,volume_mm3,centroid_ras], n_labels, body_label)."""
    if hasattr(results, "table"):                     # detector.BatchResult: vectorised from the C arrays
        return results.table(TABLE_SLOTS)
    t = np.zeros((len(results), TABLE_SLOTS, TABLE_FIELDS), dtype=np.float64)
    for i, r in enumerate(results):
        for j, m in enumerate(r.markers[:TABLE_SLOTS]):
            t[i, j] = (m.label, m.count, m.volume_mm3, *m.centroid_ras, r.n_labels, r.body_label)
    return t


def gather_tables(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gathers the ranks' [S, SLOTS, FIELDS] tables into [world*S, SLOTS, FIELDS] (rank-major)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def unshard(gathered: torch.Tensor, n_items: int, world: int) -> torch.Tensor:
    """Reorders a rank-major gathered table (equal shard sizes, padded) back to item order."""
    per = gathered.shape[0] // world
    order = []
    for i in range(n_items):
        order.append((i % world) * per + i // world)
    return gathered[torch.as_tensor(order, device=gathered.device)]


def gather_entry_results(local_index: int, local_distance: float, offset: int, device, group=None):
    """Global arg-min of the per-rank entry-search winners (lowest global index on ties).  `offset` is the
    first global candidate index of this rank's block; local_index < 0 means no suitable point."""
    gi = float(local_index + offset) if local_index >= 0 else float("inf")
    me = torch.tensor([[local_distance if local_index >= 0 else float("inf"), gi]], dtype=torch.float64, device=device)
    allr = gather_tables(me, group)
    best, best_d = -1, float("inf")
    for d, i in allr.tolist():
        if i != float("inf") and (d < best_d or (d == best_d and i < best)):
            best, best_d = int(i), d
    return best, best_d
