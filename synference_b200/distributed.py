"""One process per GPU: sharding rule and the (only) collective of the path.

The reference shards galaxies contiguously over MPI ranks and issues no collective of its own
(``library.py:3127-3138``; rank files are merged on disk, ``utils.py:2214-2328``).  Here ranks come
from ``torch.distributed`` (NCCL over NVLink on B200, gloo in CPU tests): every rank synthesises its
slice and writes its own shard; ``gather_rows`` all-gathers the optional in-memory ``(N, n_feat)``
training tensor.  Nothing else on the data path communicates.
"""

from __future__ import annotations

import os

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def rank_world():
    d = _dist()
    if d is not None:
        return d.get_rank(), d.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def local_device() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(total: int, rank: int, size: int):
    """Rows ``[rank*(total//size), ...)``; the last rank takes the remainder (``library.py:3130-3137``)."""
    per = total // size
    start = rank * per
    end = total if rank == size - 1 else start + per
    return start, end


def shard_counts(total: int, size: int):
    return [shard_bounds(total, r, size)[1] - shard_bounds(total, r, size)[0] for r in range(size)]


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def gather_rows(local, total: int):
    """All-gather row blocks of unequal length into the full ``(total, n_feat)`` tensor on every rank.

    ``local`` is this rank's ``(n_local, n_feat)`` torch tensor (CUDA with NCCL, CPU with gloo).
    Shards are padded to the largest count so a single ``all_gather_into_tensor`` suffices.
    """
    import torch
    d = _dist()
    if d is None:
        return local
    rank, size = d.get_rank(), d.get_world_size()
    counts = shard_counts(total, size)
    assert local.shape[0] == counts[rank], f"rank {rank}: expected {counts[rank]} rows, got {local.shape[0]}"
    width = max(counts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((size * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    try:
        d.all_gather_into_tensor(out, pad)
    except (RuntimeError, NotImplementedError):  # backends without the fused form
        parts = [torch.empty_like(pad) for _ in range(size)]
        d.all_gather(parts, pad)
        out = torch.cat(parts, 0)
    out = out.view((size, width) + tuple(local.shape[1:]))
    return torch.cat([out[r, : counts[r]] for r in range(size)], 0)


def merge_rank_shards(paths, out_path):
    """Host-side merge of per-rank LIBRARY shards (``Grid/Photometry`` ... with galaxies along the LAST axis, as
    ``CombinedBasis.save_library`` writes them) into one library; everything that is not per-galaxy (the ``Model`` group)
    is taken from the first shard."""
    from .utils import read_container, write_container
    datasets, attrs = None, None
    for p in paths:
        d, a = read_container(p)
        if datasets is None:
            datasets, attrs = {k: [v] for k, v in d.items()}, dict(a)
        else:
            for k, v in d.items():
                if k.startswith("Grid/"):
                    datasets[k].append(v)
    merged = {k: (np.concatenate(v, axis=-1) if k.startswith("Grid/") else v[0]) for k, v in datasets.items()}
    attrs["world_size"], attrs["rank"] = 1, 0
    write_container(out_path, merged, attrs)
    return out_path
