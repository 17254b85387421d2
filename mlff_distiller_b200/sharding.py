"""Multi-GPU sharding of batched workloads.

Structures are independent -- the reference masks every cross-structure pair
(src/mlff_distiller/models/student_model.py:94-99) -- so a batch shards by structure with NO
collective on the data path: one process per GPU, one contiguous shard per rank, balanced by
atom count; results are concatenated on the host in input order.  A single large periodic
system or a single MD trajectory does not shard ("replicas only").
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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


_PINNED = {}   # width -> [two pinned host tensors, index of the one handed out last]


def _pinned(rows: int, width: int):
    """Page-locked staging buffer for the root of gather_in_order.  Page-locking 60 MB costs ~10 ms, so the
    buffers are kept; two per row width are handed out alternately, so the arrays a gather returns stay
    valid until the second-next gather of the same width (copy them to keep them longer)."""
    import torch
    slot = _PINNED.setdefault(width, [[None, None], 1])
    i = slot[1] = 1 - slot[1]
    buf = slot[0][i]
    if buf is None or buf.shape[0] < rows:
        buf = slot[0][i] = torch.empty((max(rows, 1), width), dtype=torch.float32).pin_memory()
    return buf[:rows]


def gather_in_order(local_energies: np.ndarray, local_forces: np.ndarray, counts: Sequence[int],
                    group=None, device=None, root: Optional[int] = None):
    """Gather per-rank results into input order with fixed-size transfers and no pickling.

    ``counts`` is the global per-structure atom count list every rank already holds (the shards are a
    pure function of it, :func:`partition_by_atoms`), so every rank knows every shard's size and its
    place in the output.  Works with gloo (CPU tensors, tests) and with nccl (``device`` = this rank's
    CUDA device).

    ``root=None``: every rank receives everything (two padded ``all_gather_into_tensor`` calls).
    ``root=r``: shards are contiguous in input order, so every other rank sends its slab point-to-point
    straight into its slice of ONE output tensor on rank r (no padding, no concatenation; one
    device-to-pinned-host copy); rank r returns ``(energies, forces)``, the others ``(None, None)``.  On CUDA
    the returned arrays are views of cached pinned buffers: valid until the second-next gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = np.asarray(counts, dtype=np.int64)
    shards = partition_by_atoms(counts, world)
    n_structs = [b - a for a, b in shards]
    n_atoms = [int(counts[a:b].sum()) for a, b in shards]
    local_energies = np.asarray(local_energies, dtype=np.float32).reshape(-1)
    local_forces = np.asarray(local_forces, dtype=np.float32).reshape(-1, 3)
    want_forces = len(local_forces) > 0 or n_atoms[rank] == 0
    if len(local_energies) != n_structs[rank] or (want_forces and len(local_forces) != n_atoms[rank]):
        raise ValueError("local results do not match this rank's shard of `counts`")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    device = torch.device(device)

    if root is not None:
        def to_root(local: np.ndarray, sizes, width: int):
            starts = np.concatenate([[0], np.cumsum(sizes)])
            mine = torch.from_numpy(np.ascontiguousarray(local).reshape(-1, width)).to(device)
            if rank != root:
                if sizes[rank]:
                    dist.send(mine, dst=root, group=group)
                return None
            out = torch.empty((int(starts[-1]), width), dtype=torch.float32, device=device)
            out[int(starts[root]):int(starts[root + 1])] = mine
            for r in range(world):
                if r != root and sizes[r]:
                    dist.recv(out[int(starts[r]):int(starts[r + 1])], src=r, group=group)
            if device.type == "cuda":
                host = _pinned(out.shape[0], width)      # page-locked once, reused (alternating) by later gathers
                host.copy_(out, non_blocking=False)
                return host.numpy()
            return out.numpy()

        energies = to_root(local_energies, n_structs, 1)
        forces = to_root(local_forces, n_atoms, 3) if want_forces else None
        if rank != root:
            return None, None
        return energies.reshape(-1), (forces if want_forces else np.zeros((0, 3), dtype=np.float32))

    def exchange(local: np.ndarray, sizes, width: int) -> np.ndarray:
        cap = max(max(sizes), 1)
        send = torch.zeros((cap, width), dtype=torch.float32, device=device)
        if len(local):
            send[: len(local)] = torch.from_numpy(np.ascontiguousarray(local).reshape(-1, width)).to(device)
        recv = torch.empty((world * cap, width), dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(recv, send, group=group)
        host = recv.cpu().numpy().reshape(world, cap, width)
        return np.concatenate([host[r, : sizes[r]] for r in range(world)])

    energies = exchange(local_energies, n_structs, 1).reshape(-1)
    forces = exchange(local_forces, n_atoms, 3) if want_forces else np.zeros((0, 3), dtype=np.float32)
    return energies, forces


class SharedResults:
    """Ordered gather of per-rank HOST results on one node through POSIX shared memory.

    The batched interface hands every rank its energies and forces as host arrays (the device-to-host
    copies are pipelined under the kernels, ``StudentForceFieldCalculator.evaluate_stream``).  On one
    node -- the scope of ``bench.py --gpus N`` and of the reference's single-host sweeps -- the cheapest
    ordered gather is therefore no transfer at all: the root creates one shared-memory segment laid out
    in input order (``[B]`` energies, ``[N, 3]`` forces, float32), every rank maps it and copies its
    shard to its own offset (all ranks in parallel, ~1 ms for 8 MB), and one barrier later the root reads
    the complete arrays.  No pickling, no padding, no host -> device -> host round trip of results.
    ``torch.distributed`` (any backend) is used for the segment name (once) and the barrier only.
    Multi-node jobs use :func:`gather_in_order`.
    """

    def __init__(self, counts: Sequence[int], group=None, root: int = 0):
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self._dist, self.group, self.root = dist, group, root
        self.counts = np.asarray(counts, dtype=np.int64)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.a, self.b = shard_slice(self.counts, self.rank, world)
        prefix = np.concatenate([[0], np.cumsum(self.counts)])
        self.atom0, self.atom1 = int(prefix[self.a]), int(prefix[self.b])
        nb, na = len(self.counts), int(prefix[-1])
        nbytes = 4 * nb + 12 * na
        name = [None]
        if self.rank == root:
            self._shm = shared_memory.SharedMemory(create=True, size=max(nbytes, 16))
            name[0] = self._shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=root, group=group)   # once, outside any timed region
        if self.rank != root:
            self._shm = shared_memory.SharedMemory(name=name[0])
            # Python < 3.13 registers an ATTACHED segment with this process's resource tracker too, which would
            # unlink it when this rank exits (possibly before the root has read it) and warn about a leak; the
            # root created the segment and is the one that unlinks it
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.energies = np.ndarray((nb,), dtype=np.float32, buffer=self._shm.buf, offset=0)
        self.forces = np.ndarray((na, 3), dtype=np.float32, buffer=self._shm.buf, offset=4 * nb)

    def write(self, local_energies: np.ndarray, local_forces: Optional[np.ndarray] = None):
        """Copy this rank's shard (structures [a, b) of the global list) to its place."""
        self.energies[self.a:self.b] = np.asarray(local_energies, dtype=np.float32).reshape(-1)
        if local_forces is not None and len(local_forces):
            self.forces[self.atom0:self.atom1] = np.asarray(local_forces, dtype=np.float32).reshape(-1, 3)

    def collect(self):
        """Barrier, then ``(energies, forces)`` views on the root and ``(None, None)`` elsewhere.  The views
        stay valid until :meth:`close`; the next :meth:`write` of any rank overwrites them."""
        if self._dist.is_initialized() and self._dist.get_world_size(self.group) > 1:
            self._dist.barrier(group=self.group)
        if self.rank == self.root:
            return self.energies, self.forces
        return None, None

    def close(self):
        if self._shm is None:
            return
        if self._dist.is_initialized() and self._dist.get_world_size(self.group) > 1:
            self._dist.barrier(group=self.group)   # nobody is still writing
        self.energies = self.forces = None
        try:
            self._shm.close()
        except BufferError:   # the caller still holds views of the arrays; the mapping goes with them
            pass
        if self.rank == self.root:
            self._shm.unlink()
        self._shm = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

