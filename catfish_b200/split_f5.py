"""Signal slicing after chunk merging ("next" row N2 of SURVEY.md section 8f).

Mirror of /root/reference/catfish/split_f5.py (``split_signal`` :8-81).  The only computation in
the reference function is ``signal_dset[s[0] : s[1]]`` for every homopolymer / non-homopolymer
range; that runs batched on the GPU (K8, ``cf_split_raw``).  Rewriting FAST5 containers (copy the
file, replace the ``Signal`` dataset with the gzip-9 int16 piece, set ``duration``, drop earlier
basecalls) is HDF5 I/O and needs h5py, which is guarded like in ``catfish_b200.infer``.
"""

import os

import numpy as np

from . import _cabi


def split_raw(raws, ranges_per_read, device=None):
    """``[[raw[s:e] for (s, e) in ranges] for raw, ranges in zip(raws, ranges_per_read)]`` on the GPU.

    ``raws``: list of int16 arrays; ``ranges_per_read``: per read a list of (start, end) pairs (end
    exclusive, numpy slice semantics).  Returns per read the list of int16 pieces."""
    import torch
    from . import get_device
    dev = get_device() if device is None else int(device)
    arrays = [np.ascontiguousarray(np.asarray(r).reshape(-1), dtype=np.int16) for r in raws]
    offsets = np.zeros(len(arrays) + 1, np.int64)
    if arrays:
        offsets[1:] = np.cumsum([a.size for a in arrays])
    flat = np.concatenate(arrays) if arrays else np.zeros(0, np.int16)
    rng, owner = [], []
    for r, ranges in enumerate(ranges_per_read):
        for s, e in ranges:
            rng.append((int(s), int(e)))
            owner.append(r)
    n = len(rng)
    if n == 0:
        return [[] for _ in arrays]
    capacity = int(sum(arrays[o].size for o in owner))          # a piece is never longer than its read
    lib = _cabi.load_library()
    with torch.cuda.device(dev):
        d = "cuda:%d" % dev
        raw_d = torch.from_numpy(flat).to(d) if flat.size else torch.zeros(1, dtype=torch.int16, device=d)
        rng_d = torch.tensor(rng, dtype=torch.int64, device=d).reshape(-1, 2)
        own_d = torch.tensor(owner, dtype=torch.int32, device=d)
        out_d = torch.empty(max(capacity, 1), dtype=torch.int16, device=d)
        off_d = torch.zeros(n + 1, dtype=torch.int64, device=d)
        _cabi.check(lib.cf_split_raw(dev, raw_d.data_ptr(), offsets.ctypes.data_as(_cabi.c_i64_p), len(arrays),
                                     rng_d.data_ptr(), own_d.data_ptr(), n, out_d.data_ptr(), capacity,
                                     off_d.data_ptr(), torch.cuda.current_stream().cuda_stream))
        off = off_d.cpu().numpy()
        out = out_d[:int(off[-1])].cpu().numpy()
    pieces = [[] for _ in arrays]
    for i, o in enumerate(owner):
        pieces[o].append(out[off[i]:off[i + 1]].copy())
    return pieces


def split_signal(input_file, splits_hp, splits_nonhp, temp_dir, temp_dir_nonhp):
    """split_f5.py:8-81: write one FAST5 per range (needs h5py).  Same arguments and return value
    (two lists with the read name repeated once per written file)."""
    import h5py                                         # not a dependency of the array-level API
    from shutil import copyfile
    try:
        source = h5py.File(input_file, "r")
    except IOError:
        raise IOError("ERROR - could not open file, likely corrupted.")
    try:
        read_name = list(source["Raw"]["Reads"])[0]
        signal = source["Raw"]["Reads"][read_name]["Signal"][()]
    except Exception:
        raise RuntimeError("ERROR - no raw signal data was stored in file.")
    ranges = [tuple(s[:2]) for s in splits_hp] + [tuple(s[:2]) for s in splits_nonhp]
    pieces = split_raw([signal], [ranges])[0]
    stem = os.path.basename(input_file).split(".")[0]
    written = ([], [])
    for index, piece in enumerate(pieces):
        is_hp = index < len(splits_hp)
        dest_name = "{}/{}_{}.fast5".format(temp_dir if is_hp else temp_dir_nonhp, stem, index)
        copyfile(input_file, dest_name)
        with h5py.File(dest_name, "r+") as dest:
            reads = dest["Raw"]["Reads"][read_name]
            del reads["Signal"]
            reads.create_dataset("Signal", data=piece, dtype="int16", compression="gzip", compression_opts=9)
            if is_hp:
                dest["Raw"]["Reads"]["duration"] = len(piece)
                for group in ("Basecall_1D_000", "RawGenomeCorrected_000"):
                    if dest["Analyses"][group]:
                        del dest["Analyses"][group]
        written[0 if is_hp else 1].append(read_name)
    return written
