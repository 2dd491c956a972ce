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


def gather_intervals(local_indices, local_hps, local_lengths, n_reads, rank, world_size, dst=0):
    """Host-side gather of per-read interval lists in compact form: every rank sends ONE tuple of numpy arrays
    (read indices, read lengths, CSR offsets, intervals [n, 2]) instead of one Python object per interval, and
    rank ``dst`` rebuilds the per-read ``IntervalList`` views.  Returns (hps, lengths) in read order on ``dst``."""
    from .infer import IntervalList
    counts = np.array([len(h) for h in local_hps], np.int64)
    ioff = np.zeros(len(local_hps) + 1, np.int64)
    np.cumsum(counts, out=ioff[1:])
    if len(local_hps):
        flat = np.concatenate([np.asarray(getattr(h, "array", h), np.int64).reshape(-1, 2) for h in local_hps])
    else:
        flat = np.zeros((0, 2), np.int64)
    payload = (np.asarray(local_indices, np.int64), np.asarray(local_lengths, np.int64), ioff, flat)
    if world_size == 1:
        gathered = [payload]
    else:
        import torch.distributed as dist
        gathered = [None] * world_size if rank == dst else None
        dist.gather_object(payload, gathered, dst=dst)
        if rank != dst:
            return None
    hps = [None] * n_reads
    lengths = [None] * n_reads
    for idx, lens, off, iv in gathered:
        bounds = off.tolist()
        for k, i in enumerate(idx.tolist()):
            hps[i] = IntervalList(iv[bounds[k]:bounds[k + 1]])
            lengths[i] = int(lens[k])
    return hps, lengths


def infer_reads_sharded(raws, model, rank, world_size, batch_reads=512, lengths=None, **kwargs):
    """The reference's loop over all reads (catfish/catfish:55-56) sharded by read over ``world_size`` GPUs:
    this rank runs ``infer.infer_reads`` over its LPT shard in batches of ``batch_reads`` reads; (hps, lengths)
    for ALL reads, in read order, on rank 0 (None elsewhere).  With ``lengths`` given, ``raws[i]`` is only touched
    for reads of this rank's shard, so a caller may pass a lazy sequence that holds just those."""
    from . import infer
    idx = shard_for_rank(lengths if lengths is not None else [len(r) for r in raws], rank, world_size)
    hps, lengths = [], []
    for b in range(0, len(idx), batch_reads):
        h, l = infer.infer_reads([raws[int(i)] for i in idx[b:b + batch_reads]], model, **kwargs)
        hps.extend(h)
        lengths.extend(l)
    return gather_intervals(idx, hps, lengths, len(raws), rank, world_size)
