"""Sharding reads over the GPUs of one box.

Reads are independent (the reference loops over files with no cross-read state,
/root/reference/catfish/catfish:55-56), so the path shards by read ID with no
collective on the compute path: every rank runs the full kernel sequence on its
own reads; only the per-read results (intervals, read length) are gathered on the
host of rank 0 (``torch.distributed.gather_object``, any backend).
"""

import heapq

import numpy as np


def partition_reads(lengths, n_parts):
    """Longest-processing-time greedy assignment of reads to ``n_parts`` shards by sample count.

    Returns a list of ``n_parts`` index arrays (each sorted ascending); deterministic."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    heap = [(0, p) for p in range(n_parts)]
    heapq.heapify(heap)
    parts = [[] for _ in range(n_parts)]
    for i in order:
        load, p = heapq.heappop(heap)
        parts[p].append(int(i))
        heapq.heappush(heap, (load + int(lengths[i]), p))
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def shard_for_rank(lengths, rank, world_size):
    return partition_reads(lengths, world_size)[rank]


def gather_results(local_indices, local_results, n_reads, rank, world_size, dst=0):
    """Host-side gather of per-read results keyed by read index.

    ``local_results[i]`` belongs to read ``local_indices[i]``.  Returns the list of all
    ``n_reads`` results in read order on rank ``dst`` (None elsewhere)."""
    if world_size == 1:
        out = [None] * n_reads
        for i, res in zip(local_indices, local_results):
            out[int(i)] = res
        return out
    import torch.distributed as dist
    payload = (np.asarray(local_indices).tolist(), list(local_results))
    gathered = [None] * world_size if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    out = [None] * n_reads
    for idx, res in gathered:
        for i, r in zip(idx, res):
            out[int(i)] = r
    return out


def infer_reads_sharded(raws, model, rank, world_size, **kwargs):
    """``infer.infer_reads`` over this rank's shard; (hps, lengths) for all reads on rank 0."""
    from . import infer
    idx = shard_for_rank([len(r) for r in raws], rank, world_size)
    hps, lengths = infer.infer_reads([raws[int(i)] for i in idx], model, **kwargs) if len(idx) else ([], [])
    merged = gather_results(idx, list(zip(hps, lengths)), len(raws), rank, world_size)
    if merged is None:
        return None
    return [m[0] for m in merged], [m[1] for m in merged]
