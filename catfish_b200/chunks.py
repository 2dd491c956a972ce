"""Chunk merging after inference ("next" row N1 of SURVEY.md section 8f).

Mirror of the per-file loop body of the reference's CLI, /root/reference/catfish/catfish:58-81, and of
``center_hp`` (:121-135): homopolymer intervals of a read are merged into chunks of at least
``chunk_size`` samples, short chunks are centre-padded and clamped to the read, and the complement is
listed as non-homopolymer ranges.  The arithmetic runs in the K7 kernel (``cf_merge_chunks``), one
thread per read, reproducing the reference's list aliasing bit for bit.
"""

import numpy as np

from . import _cabi


def center_hp(merged_positions, len_read, chunk_size=1000):
    """catfish:121-135: pad the LAST chunk symmetrically up to ``chunk_size`` and keep it inside the
    read (in place, host helper - one interval of integer arithmetic; the batched path is K7)."""
    chunk = merged_positions[-1]
    missing = chunk_size - (chunk[1] - chunk[0])
    if missing > 0:
        chunk[0] -= missing // 2
        chunk[1] += missing - missing // 2
        if chunk[0] < 0:                       # shift right so that the chunk starts at 0
            chunk[1] -= chunk[0]
            chunk[0] = 0
        if chunk[1] > len_read:                # the reference moves the START by the overflow here
            chunk[0] -= len_read - chunk[1]
            chunk[1] = len_read
    return merged_positions


def merge_reads(hp_positions_per_read, read_lengths, chunk_size=1000, device=None):
    """Batched catfish:58-81.  Returns ``(hp_chunks, nonhp_ranges)``: per read the merged positions
    (list of [start, end]) and the non-HP ranges in exactly the structures the reference stores in
    ``hp_dict[file]`` / ``nonhp_dict[file]`` (a read without intervals has no hp entry -> ``None`` -
    and the reference's ``[([(0, len_read), len_read])]`` non-HP entry)."""
    import torch
    from . import get_device
    dev = get_device() if device is None else int(device)
    n_reads = len(hp_positions_per_read)
    counts = [len(h) for h in hp_positions_per_read]
    ioff = np.zeros(n_reads + 1, np.int64)
    ioff[1:] = np.cumsum(counts)
    n_int = int(ioff[-1])
    flat = np.zeros((max(n_int, 1), 2), np.int64)
    if n_int:
        flat[:n_int] = np.concatenate([np.asarray(h, np.int64).reshape(-1, 2) for h in hp_positions_per_read if len(h)])
    lengths = np.ascontiguousarray(read_lengths, dtype=np.int64)
    lib = _cabi.load_library()
    with torch.cuda.device(dev):
        d = "cuda:%d" % dev
        iv = torch.from_numpy(flat).to(d)
        merged = torch.zeros((n_int + n_reads + 1, 2), dtype=torch.int64, device=d)
        nonhp = torch.zeros((n_int + 2 * n_reads + 1, 2), dtype=torch.int64, device=d)
        mcnt = torch.zeros(max(n_reads, 1), dtype=torch.int64, device=d)
        ncnt = torch.zeros(max(n_reads, 1), dtype=torch.int64, device=d)
        _cabi.check(lib.cf_merge_chunks(dev, iv.data_ptr(), ioff.ctypes.data_as(_cabi.c_i64_p),
                                        lengths.ctypes.data_as(_cabi.c_i64_p), n_reads, int(chunk_size),
                                        merged.data_ptr(), mcnt.data_ptr(), nonhp.data_ptr(), ncnt.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
        merged, nonhp, mcnt, ncnt = merged.cpu().numpy(), nonhp.cpu().numpy(), mcnt.cpu().numpy(), ncnt.cpu().numpy()
    hp_out, non_out = [], []
    for r in range(n_reads):
        if ncnt[r] < 0:                                   # no homopolymers in this read (catfish:80-81)
            hp_out.append(None)
            non_out.append([([(0, int(lengths[r])), int(lengths[r])])])
            continue
        mo, no = int(ioff[r]) + r, int(ioff[r]) + 2 * r
        hp_out.append(merged[mo:mo + int(mcnt[r])].tolist())
        non_out.append(nonhp[no:no + int(ncnt[r])].tolist())
    return hp_out, non_out


def merge_read(hp_positions, len_read, chunk_size=1000):
    """Single-read form of ``merge_reads``: ``(merged_positions | None, nonhp_ranges)``."""
    hp, non = merge_reads([hp_positions], [len_read], chunk_size)
    return hp[0], non[0]
