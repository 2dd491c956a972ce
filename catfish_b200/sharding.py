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


_PIN = {}


def _pinned(name, n):
    import torch
    buf = _PIN.get(name)
    if buf is None or buf.numel() < n:
        buf = torch.empty(n + n // 4 + 1024, dtype=torch.int64).pin_memory()
        _PIN[name] = buf
    return buf


def reserve_gather(max_reads_per_rank, max_intervals_per_rank, rank, world_size, dst=0):
    """Pre-size the pinned staging buffers of the NCCL gather (``gather_csr``) for shards of up to the given size.
    Page-locking the receive buffer of a large job costs more than the gather itself, so a service that knows its job
    sizes reserves once; without it the buffers grow on demand and are kept for the next call."""
    need = 3 * int(max_reads_per_rank) + 1 + 2 * int(max_intervals_per_rank)
    _pinned("send", need)
    if rank == dst:
        _pinned("recv", world_size * need)


def _gather_arrays_nccl(payload, rank, world_size, dst):
    """The compact payload of ``gather_csr`` through ONE padded int64 tensor per rank over NCCL (no pickling):
    sizes by all_gather, then gather of [idx | lengths | offsets | intervals], staged through pinned host buffers.
    This is result plumbing after the compute, not a collective on the data path."""
    import torch
    import torch.distributed as dist
    idx, lens, ioff, flat = payload
    dev = torch.device("cuda", torch.cuda.current_device())
    sizes = torch.tensor([len(idx), flat.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty(2, dtype=torch.int64, device=dev) for _ in range(world_size)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).tolist()
    need = max(3 * n + 1 + 2 * m for n, m in all_sizes)
    n, m = len(idx), flat.shape[0]
    stage = _pinned("send", need).numpy()
    np.concatenate([idx, lens, ioff, flat.reshape(-1)], out=stage[:3 * n + 1 + 2 * m])
    buf = torch.empty(need, dtype=torch.int64, device=dev)
    buf.copy_(_PIN["send"][:need], non_blocking=True)
    parts = torch.empty((world_size, need), dtype=torch.int64, device=dev) if rank == dst else None
    dist.gather(buf, list(parts.unbind(0)) if rank == dst else None, dst=dst)
    if rank != dst:
        torch.cuda.current_stream().synchronize()          # the pinned staging buffer is reused by the next call
        return None
    host = _pinned("recv", world_size * need)[:world_size * need].view(world_size, need)
    host.copy_(parts, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    # The pinned receive buffer is reused by the next call, so the parts are copied out - by a few threads: one memcpy
    # of the ~1 GB a 100 000-read job gathers was a third of the gather (numpy releases the GIL in large copies).
    out, jobs = [], []
    for r, (n, m) in enumerate(all_sizes):
        a = host[r].numpy()
        iv = np.empty((m, 2), np.int64)
        out.append((a[:n].copy(), a[n:2 * n].copy(), a[2 * n:3 * n + 1].copy(), iv))
        src = a[3 * n + 1:3 * n + 1 + 2 * m].reshape(m, 2)
        step = max(1, 1 << 21)                                  # 32 MB slices
        jobs += [(iv[k:k + step], src[k:k + step]) for k in range(0, m, step)]
    if len(jobs) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(lambda j: np.copyto(j[0], j[1]), jobs))
    else:
        for dst, src in jobs:
            np.copyto(dst, src)
    return out


class ShardedIntervals(object):
    """Sequence of per-read ``IntervalList`` over the CSR parts the ranks sent: read r lives in part ``part[r]`` at
    row ``row[r]``.  Built with vectorised numpy only (a job of 1e5 reads must not spend its time creating Python
    objects on rank 0); the per-read views are made on access."""

    def __init__(self, n_reads, parts):
        self.parts = [(off, iv) for _, _, off, iv in parts]
        self.part = np.zeros(n_reads, np.int32)
        self.row = np.zeros(n_reads, np.int64)
        for p, (idx, _, _, _) in enumerate(parts):
            self.part[idx] = p
            self.row[idx] = np.arange(len(idx))

    def __len__(self):
        return len(self.part)

    def __getitem__(self, r):
        from .infer import IntervalList
        if isinstance(r, slice):
            return [self[i] for i in range(*r.indices(len(self)))]
        off, iv = self.parts[self.part[r]]
        k = self.row[r]
        return IntervalList(iv[off[k]:off[k + 1]])

    def __iter__(self):
        return (self[r] for r in range(len(self)))

    def __eq__(self, other):
        return len(self) == len(other) and all(a == b for a, b in zip(self, other))


def gather_intervals(local_indices, local_hps, local_lengths, n_reads, rank, world_size, dst=0):
    """Host-side gather of per-read interval lists (``IntervalList`` or [n, 2] arrays): packs them into CSR
    form and calls ``gather_csr``.  Returns (hps, lengths) in read order on ``dst`` (None elsewhere)."""
    counts = np.array([len(h) for h in local_hps], np.int64)
    ioff = np.zeros(len(local_hps) + 1, np.int64)
    np.cumsum(counts, out=ioff[1:])
    if len(local_hps):
        flat = np.concatenate([np.asarray(getattr(h, "array", h), np.int64).reshape(-1, 2) for h in local_hps])
    else:
        flat = np.zeros((0, 2), np.int64)
    return gather_csr(local_indices, local_lengths, ioff, flat, n_reads, rank, world_size, dst)


def gather_csr(local_indices, local_lengths, ioff, flat, n_reads, rank, world_size, dst=0):
    """Every rank sends ONE tuple of numpy arrays (read indices, read lengths, CSR offsets, intervals [n, 2]) instead
    of one Python object per interval; rank ``dst`` returns (``ShardedIntervals``, lengths) in read order."""
    payload = (np.asarray(local_indices, np.int64), np.asarray(local_lengths, np.int64), np.asarray(ioff, np.int64),
               np.asarray(flat, np.int64).reshape(-1, 2))
    if world_size == 1:
        gathered = [payload]
    else:
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            gathered = _gather_arrays_nccl(payload, rank, world_size, dst)
        else:
            gathered = [None] * world_size if rank == dst else None
            dist.gather_object(payload, gathered, dst=dst)
        if rank != dst:
            return None
    lengths = np.zeros(n_reads, np.int64)
    for idx, lens, _, _ in gathered:
        lengths[idx] = lens
    return ShardedIntervals(n_reads, gathered), lengths.tolist()


def infer_reads_sharded(raws, model, rank, world_size, batch_reads=512, lengths=None, **kwargs):
    """The reference's loop over all reads (catfish/catfish:55-56) sharded by read over ``world_size`` GPUs:
    this rank runs ``infer.infer_reads_arrays`` over its LPT shard in batches of ``batch_reads`` reads; (hps, lengths)
    for ALL reads, in read order, on rank 0 (None elsewhere).  With ``lengths`` given, ``raws[i]`` is only touched
    for reads of this rank's shard, so a caller may pass a lazy sequence that holds just those."""
    import time
    from . import infer
    t0 = time.perf_counter()
    idx = shard_for_rank(lengths if lengths is not None else [len(r) for r in raws], rank, world_size)
    t1 = time.perf_counter()
    flats, offs, lens, base = [], [np.zeros(1, np.int64)], [], 0
    for b in range(0, len(idx), batch_reads):
        iv, ioff, ln = infer.infer_reads_arrays([raws[int(i)] for i in idx[b:b + batch_reads]], model, **kwargs)[:3]
        flats.append(iv)
        offs.append(ioff[1:] + base)
        lens.append(ln)
        base += int(ioff[-1])
    t2 = time.perf_counter()
    flat = np.concatenate(flats) if flats else np.zeros((0, 2), np.int64)
    out = gather_csr(idx, np.concatenate(lens) if lens else np.zeros(0, np.int64), np.concatenate(offs), flat,
                     len(raws), rank, world_size)
    last_timing.update(partition_s=t1 - t0, infer_s=t2 - t1, gather_s=time.perf_counter() - t2)
    return out


last_timing = {}          # host wall-clock breakdown of this rank's last infer_reads_sharded call (bench.py reports it)
