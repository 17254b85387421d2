"""Multi-GPU sharding of batched workloads.

Structures are independent -- the reference masks every cross-structure pair
(src/mlff_distiller/models/student_model.py:94-99) -- so a batch shards by structure with NO
collective on the data path: one process per GPU, one contiguous shard per rank, balanced by
atom count; results are concatenated on the host in input order.  A single large periodic
system or a single MD trajectory does not shard ("replicas only").
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def partition_by_atoms(counts: Sequence[int], num_shards: int) -> List[Tuple[int, int]]:
    """Contiguous ``[start, end)`` structure ranges whose atom totals are as even as a
    prefix-sum split allows.  Every structure lands in exactly one shard; shards may be empty
    when there are fewer structures than shards."""
    counts = np.asarray(counts, dtype=np.int64)
    n = len(counts)
    if num_shards < 1:
        raise ValueError("num_shards must be >= 1")
    prefix = np.concatenate([[0], np.cumsum(counts)])
    total = int(prefix[-1])
    bounds = [0]
    for r in range(1, num_shards):
        target = total * r / num_shards
        cut = int(np.searchsorted(prefix, target, side="left"))
        # choose the neighbour cut that is closer to the target
        if cut > 0 and abs(prefix[cut - 1] - target) <= abs(prefix[min(cut, n)] - target):
            cut -= 1
        cut = min(max(cut, bounds[-1]), n)
        bounds.append(cut)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(num_shards)]


def shard_slice(counts: Sequence[int], rank: int, world_size: int) -> Tuple[int, int]:
    return partition_by_atoms(counts, world_size)[rank]


def chunk_by_budget(counts: Sequence[int], max_atoms: int, max_structs: int) -> List[Tuple[int, int]]:
    """Split a shard into micro-batches that fit the workspace (atoms and structures)."""
    out, start, atoms = [], 0, 0
    for i, c in enumerate(counts):
        if i > start and (atoms + c > max_atoms or i - start >= max_structs):
            out.append((start, i))
            start, atoms = i, 0
        atoms += int(c)
    if start < len(counts):
        out.append((start, len(counts)))
    return out


def gather_in_order(local_energies: np.ndarray, local_forces: np.ndarray, group=None):
    """Host-side gather of per-rank results into input order via ``all_gather_object``
    (works on gloo for CPU tests and on nccl-initialised jobs alike; results are small:
    4 B per structure + 12 B per atom)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, (np.asarray(local_energies), np.asarray(local_forces)), group=group)
    energies = np.concatenate([b[0] for b in bucket])
    forces = np.concatenate([b[1] for b in bucket])
    return energies, forces
