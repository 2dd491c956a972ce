"""CPU: the N>1 host path (shard by read, host-side gather) on world_size-2 gloo."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from catfish_b200 import sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_infer(raw):
    """Stand-in for the per-read GPU result: depends only on the read's content."""
    return [[int(raw[0]), int(raw.sum() % 1000)]], len(raw)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = synth.ragged_lengths(37, 100, 2000, seed=1)
    reads = synth.synth_reads(lengths, base_seed=50)
    idx = sharding.shard_for_rank(lengths, rank, world)
    local = [_fake_infer(reads[int(i)]) for i in idx]
    merged = sharding.gather_results(idx, local, len(reads), rank, world)
    # the compact path the sharded job uses: one tuple of arrays per rank, IntervalList views on rank 0
    hps = [np.array([[int(r[0]), k] for k in range(int(r[1]) % 4)], np.int64).reshape(-1, 2) for r in (reads[int(i)] for i in idx)]
    compact = sharding.gather_intervals(idx, hps, [len(reads[int(i)]) for i in idx], len(reads), rank, world)
    if rank == 0:
        want_hps = [[[int(r[0]), k] for k in range(int(r[1]) % 4)] for r in reads]
        assert [h.tolist() for h in compact[0]] == want_hps and compact[1] == [len(r) for r in reads]
        assert all(h == w for h, w in zip(compact[0], want_hps))          # list-like equality of IntervalList
        q.put((merged, [_fake_infer(r) for r in reads], [int(lengths[p].sum()) for p in sharding.partition_reads(lengths, world)]))
    else:
        assert merged is None and compact is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, want, loads = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert merged == want
    assert abs(loads[0] - loads[1]) <= 2000
