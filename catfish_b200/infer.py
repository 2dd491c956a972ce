"""Mirror of /root/reference/catfish/infer.py: the per-read inference driver.

Same function names, arguments, defaults and return types as the reference;
the arithmetic runs on the GPU through the C ABI (include/catfish_b200.h):

===========================  ==========================  =========================
reference (catfish/infer.py)  here                        C-ABI entry
===========================  ==========================  =========================
infer_class_from_signal :12   infer_class_from_signal     cf_infer_reads_host
(array-level twins)           infer_class_from_raw,       cf_infer_reads_host
                              infer_reads
process_signal :77            process_signal (h5py)       cf_normalize_reads
normalize_raw_signal :96      normalize_raw_signal        cf_normalize_reads
reshape_input :108            reshape_input               (pure reshape, host)
class_from_threshold :128     class_from_threshold        cf_class_from_threshold
hp_in_pred :141               hp_in_pred                  cf_hp_in_pred
correct_short :174            correct_short               cf_correct_short
===========================  ==========================  =========================

PyTorch tensors are used only to hold device buffers and to name the stream.
"""

import collections.abc
import ctypes
import os

import numpy as np

from . import _cabi

_STAGING = {}        # device index -> pinned int16 torch tensor the ragged batch is concatenated into


class IntervalList(collections.abc.Sequence):
    """The ``[[start, end], ...]`` list of one read (infer.py:149-162), backed by the int64 [n, 2] slice of the
    batch result.  Behaves like the reference's list of two-element lists of Python ints (indexing, iteration,
    ``len``, ``==`` with a list), but the Python objects are only built when asked for: a 512-read batch holds
    ~3e5 intervals and eager conversion costs more than the GPU work of the whole batch."""

    __slots__ = ("array",)

    def __init__(self, array):
        self.array = array

    def __len__(self):
        return int(self.array.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.array[i].tolist()
        return self.array[i].tolist()

    def __iter__(self):
        return iter(self.array.tolist())

    def __eq__(self, other):
        if isinstance(other, IntervalList):
            return np.array_equal(self.array, other.array)
        if isinstance(other, (list, tuple)):
            return self.array.tolist() == [list(x) if isinstance(x, tuple) else x for x in other]
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def tolist(self):
        return self.array.tolist()

    def __repr__(self):
        return repr(self.array.tolist())

    def __reduce__(self):
        return (IntervalList, (np.ascontiguousarray(self.array),))


_POOL = None


def _concat_into(stage, arrays, offsets):
    """Ragged concat of the batch into the pinned staging buffer.  Large batches are split into a few
    contiguous groups copied by worker threads (numpy releases the GIL while copying): one thread moves
    ~10 GB/s, which for a 128 MB batch is a tenth of the GPU time of the whole call."""
    global _POOL
    total = int(offsets[-1])
    workers = min(8, os.cpu_count() or 1)
    if total < (1 << 23) or len(arrays) < 2 * workers or workers < 2:
        np.concatenate(arrays, out=stage)
        return
    if _POOL is None:
        import concurrent.futures
        _POOL = concurrent.futures.ThreadPoolExecutor(max_workers=workers, thread_name_prefix="catfish-stage")
    cuts = np.searchsorted(offsets, np.linspace(0, total, workers + 1)[1:-1]).tolist()
    bounds = [0] + cuts + [len(arrays)]
    jobs = []
    for lo, hi in zip(bounds, bounds[1:]):
        if hi > lo:
            jobs.append(_POOL.submit(np.concatenate, arrays[lo:hi], out=stage[int(offsets[lo]):int(offsets[hi])]))
    for j in jobs:
        j.result()


def _staging(torch, device, n):
    """Pinned host buffer of at least n int16 samples for `device` (grows geometrically, reused across calls)."""
    buf = _STAGING.get(device)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n + n // 4, 1 << 20), dtype=torch.int16).pin_memory()
        _STAGING[device] = buf
    return buf


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _cabi.CatfishError("catfish_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _device_index(device=None):
    from . import get_device
    return get_device() if device is None else int(device)


def _offsets_ptr(offsets):
    return offsets.ctypes.data_as(_cabi.c_i64_p)


# ------------------------------------------------------------------------------ inference
def infer_class_from_signal(fast5_file, model, label=1, window_size=35):
    """infer.py:12-51: FAST5 path -> (list of [start, end] homopolymer intervals, read length)."""
    if not os.path.exists(fast5_file):
        raise ValueError("path to FAST5 is not correct.")
    import h5py                                   # not a dependency of the array-level API
    with h5py.File(fast5_file, "r") as fast5:
        raw = _trimmed_raw(fast5)
    return infer_class_from_raw(raw, model, label=label, window_size=window_size)


def infer_class_from_raw(raw, model, label=1, window_size=35, threshold=0.5):
    """Array-level twin of infer_class_from_signal: ``raw`` is the int16 signal of one read
    with the leading ``first_sample_template`` samples already dropped (infer.py:87-90)."""
    hps, lengths = infer_reads([raw], model, threshold=threshold, window_size=window_size)
    return hps[0].tolist(), lengths[0]          # a plain list of [int, int], exactly the reference's type


def infer_reads(raws, model, threshold=0.5, min_run=15, extension_left=11, extension_right=16,
                window_size=35, return_scores=False):
    """Batched infer_class_from_signal over a list of int16 reads (ragged).

    Returns ``(hps, lengths)`` - per read the [start, end] intervals (an ``IntervalList``: list-like, Python
    ints on access, as the reference) and ``len(labels)`` - plus the per-position float32 scores when
    ``return_scores`` is set.  One C-ABI call: host->device copy of the signal, median/MAD
    normalisation, windowing, network, threshold / short-run removal / interval emission,
    device->host copy of the results.  The ragged batch is concatenated straight into a pinned staging
    buffer (one host pass, then DMA at PCIe rate instead of a pageable copy)."""
    res = infer_reads_arrays(raws, model, threshold, min_run, extension_left, extension_right, window_size,
                             return_scores)
    intervals, ioff, lengths = res[0], res[1], res[2]
    bounds = ioff.tolist()
    hps = [IntervalList(intervals[bounds[r]:bounds[r + 1]]) for r in range(len(lengths))]
    if return_scores:
        return hps, lengths.tolist(), res[3]
    return hps, lengths.tolist()


def infer_reads_arrays(raws, model, threshold=0.5, min_run=15, extension_left=11, extension_right=16,
                       window_size=35, return_scores=False):
    """``infer_reads`` with the result left in CSR form: ``(intervals int64 [n, 2], interval_offsets int64 [R + 1],
    lengths int64 [R][, scores list])`` - what the sharded job gathers (no per-read Python objects)."""
    if window_size != model.window:
        raise ValueError("window_size must equal the model's window (%d)" % model.window)
    arrays = [_as_int16(r) for r in raws]
    for a in arrays:
        if a.size == 0:
            raise IndexError("list index out of range")        # infer.py:184 on an empty read
    n_reads = len(arrays)
    lengths = np.array([a.size for a in arrays], np.int64)
    offsets = np.zeros(n_reads + 1, np.int64)
    np.cumsum(lengths, out=offsets[1:])
    total = int(offsets[-1])
    if total >= _PIPELINE_MIN_SAMPLES and not return_scores:
        intervals, ioff = _infer_reads_pipelined(arrays, offsets, model, threshold, min_run, extension_left,
                                                 extension_right)
        return intervals, ioff, lengths
    if n_reads:
        stage = _staging(_torch(), int(model.device), total).numpy()[:total]
        _concat_into(stage, arrays, offsets)
        raw = stage
    else:
        raw = np.zeros(0, np.int16)
    res = infer_concatenated(raw, offsets, model, threshold, min_run, extension_left, extension_right,
                             return_scores)
    if return_scores:
        scores = [res[2][int(offsets[r]):int(offsets[r + 1])] for r in range(n_reads)]
        return res[0], res[1], lengths, scores
    return res[0], res[1], lengths


_PIPELINE_MIN_SAMPLES = 24_000_000      # batches of at least ~2 engine passes take the pipelined path
_GROUP_SAMPLES = 10_400_000             # one engine pass (2368 tiles of 128 windows) per group
_FIRST_GROUP_SAMPLES = int(os.environ.get("CF_FIRST_GROUP_SAMPLES", 2_600_000))
_DEVBUF = {}                            # device index -> dict of cached device / pinned result buffers


def _infer_reads_pipelined(arrays, offsets, model, threshold, min_run, ext_left, ext_right):
    """A large ragged batch as a pipeline of groups of whole reads (about one engine pass each): while the GPU
    works on group g (asynchronous ``cf_infer_reads`` on the device-resident copy), the host concatenates
    group g+1 into pinned staging and starts its copy - the host pass over the signal disappears behind the
    kernels instead of preceding them.  Same results as one ``cf_infer_reads_host`` call (reads are independent;
    batch invariance is part of the parity suite).  Returns (intervals [n, 2], interval_offsets [R + 1])."""
    torch = _torch()
    lib = _cabi.load_library()
    dev = int(model.device)
    n_reads = len(arrays)
    total = int(offsets[-1])
    # groups of consecutive reads
    cuts = [0]
    while cuts[-1] < n_reads:
        lo = cuts[-1]
        # the first group is a quarter pass: the GPU starts after ~0.5 ms of staging instead of ~2 ms
        want = _GROUP_SAMPLES if lo else _FIRST_GROUP_SAMPLES
        hi = int(np.searchsorted(offsets, offsets[lo] + want, side="right")) - 1
        cuts.append(min(n_reads, max(hi, lo + 1)))
    groups = list(zip(cuts[:-1], cuts[1:]))
    caps = [int(lib.cf_max_intervals(int(offsets[hi] - offsets[lo]), hi - lo, min_run)) for lo, hi in groups]
    cap_off = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream()
        buf = _DEVBUF.setdefault(dev, {})
        if buf.get("n_raw", 0) < total:
            buf["raw"] = torch.empty(total + total // 4, dtype=torch.int16, device="cuda")
            buf["n_raw"] = buf["raw"].numel()
        if buf.get("n_iv", 0) < cap_off[-1]:
            buf["iv"] = torch.empty((int(cap_off[-1]) * 5 // 4, 2), dtype=torch.int64, device="cuda")
            buf["iv_pin"] = torch.empty((int(cap_off[-1]) * 5 // 4, 2), dtype=torch.int64).pin_memory()
            buf["n_iv"] = buf["iv"].shape[0]
        n_off = n_reads + len(groups)
        if buf.get("n_off", 0) < n_off:
            buf["ioff"] = torch.empty(n_off * 2, dtype=torch.int64, device="cuda")
            buf["ioff_pin"] = torch.empty(n_off * 2, dtype=torch.int64).pin_memory()
            buf["n_off"] = n_off * 2
        stage_t = _staging(torch, dev, total)
        stage = stage_t.numpy()
        raw_dev, iv_dev, ioff_dev = buf["raw"], buf["iv"], buf["ioff"]
        for g, (lo, hi) in enumerate(groups):
            a, b = int(offsets[lo]), int(offsets[hi])
            _concat_into(stage[a:b], arrays[lo:hi], offsets[lo:hi + 1] - offsets[lo])
            raw_dev[a:b].copy_(stage_t[a:b], non_blocking=True)
            off_g = np.ascontiguousarray(offsets[lo:hi + 1])
            _cabi.check(lib.cf_infer_reads(
                model.handle, raw_dev.data_ptr(), _offsets_ptr(off_g), hi - lo, None,
                iv_dev.data_ptr() + 16 * int(cap_off[g]), ioff_dev.data_ptr() + 8 * (lo + g), caps[g],
                float(threshold), int(min_run), int(ext_left), int(ext_right), stream.cuda_stream))
        buf["ioff_pin"][:n_off].copy_(ioff_dev[:n_off], non_blocking=True)
        stream.synchronize()
        ioff_all = buf["ioff_pin"].numpy()
        ioff = np.zeros(n_reads + 1, np.int64)
        found = []
        for g, (lo, hi) in enumerate(groups):
            local = ioff_all[lo + g:hi + g + 1]
            n_g = int(local[-1])
            if n_g > caps[g]:
                raise _cabi.CatfishError("interval capacity exceeded (%d > %d)" % (n_g, caps[g]))
            ioff[lo + 1:hi + 1] = ioff[lo] + local[1:]
            found.append(n_g)
            if n_g:
                c0 = int(cap_off[g])
                buf["iv_pin"][c0:c0 + n_g].copy_(iv_dev[c0:c0 + n_g], non_blocking=True)
        stream.synchronize()
        iv_pin = buf["iv_pin"].numpy()
        parts = [iv_pin[int(cap_off[g]):int(cap_off[g]) + n_g] for g, n_g in enumerate(found) if n_g]
        intervals = np.concatenate(parts) if parts else np.zeros((0, 2), np.int64)
    return intervals, ioff


def infer_concatenated(raw, offsets, model, threshold=0.5, min_run=15, extension_left=11,
                       extension_right=16, return_scores=False):
    """The C-ABI call of ``infer_reads`` on an already concatenated int16 signal.

    Returns ``(intervals int64 [n,2], interval_offsets int64 [R+1][, scores float32])``."""
    torch = _torch()
    lib = _cabi.load_library()
    raw = np.ascontiguousarray(raw, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n_reads = len(offsets) - 1
    total = int(offsets[-1] - offsets[0]) if n_reads > 0 else 0
    cap = int(lib.cf_max_intervals(total, n_reads, min_run))
    intervals = np.empty((cap, 2), np.int64)
    ioff = np.zeros(n_reads + 1, np.int64)
    scores = np.empty(total, np.float32) if return_scores else None
    found = ctypes.c_int64(0)
    with torch.cuda.device(model.device):
        stream = torch.cuda.current_stream()
        _cabi.check(lib.cf_infer_reads_host(
            model.handle, raw.ctypes.data, _offsets_ptr(offsets), n_reads,
            scores.ctypes.data if scores is not None else None,
            intervals.ctypes.data, ioff.ctypes.data, cap, float(threshold), int(min_run),
            int(extension_left), int(extension_right), ctypes.byref(found), stream.cuda_stream))
    intervals = intervals[:found.value]
    if return_scores:
        return intervals, ioff, scores
    return intervals, ioff


# ------------------------------------------------------------------------------ raw signal
def _trimmed_raw(fast5_file):
    """infer.py:87-90: drop the samples before ``first_sample_template``."""
    first_sample = fast5_file["Analyses/Segmentation_000/Summary/segmentation"].attrs["first_sample_template"]
    read_name = fast5_file["Raw/Reads/"].visit(str)
    raw_signal = fast5_file["Raw/Reads/" + read_name + "/Signal"][()]
    return raw_signal[first_sample:]


def process_signal(fast5_file, normalization="median"):
    """infer.py:77-93: trimmed, normalised raw signal (float64) of an open FAST5."""
    return normalize_raw_signal(_trimmed_raw(fast5_file), normalization)


def _as_int16(raw):
    a = np.asarray(raw)
    if a.dtype == np.int16:
        return np.ascontiguousarray(a.reshape(-1))
    if a.dtype.kind in "iu" or (a.dtype.kind == "f" and np.all(a == np.rint(a))):
        if a.size and (a.min() < -32768 or a.max() > 32767):
            raise ValueError("raw signal outside the int16 DAC range")
        return np.ascontiguousarray(a.reshape(-1).astype(np.int16))
    raise ValueError("raw signal must hold integer DAC values (int16)")


def normalize_raw_signal(raw, norm_method):
    """infer.py:96-105: (raw - median) / median(|raw - median|), float64, bit-exact."""
    if norm_method != 'median':
        raise ValueError('norm_method not recognized')
    torch = _torch()
    a = _as_int16(raw)
    if a.size == 0:
        return np.zeros(0, np.float64)            # numpy: nan statistics, empty result
    dev = _device_index()
    offsets = np.array([0, a.size], np.int64)
    with torch.cuda.device(dev):
        rd = torch.from_numpy(a).to("cuda:%d" % dev)
        out = torch.empty(a.size, dtype=torch.float64, device=rd.device)
        _cabi.check(_cabi.load_library().cf_normalize_reads(
            dev, rd.data_ptr(), _offsets_ptr(offsets), 1, None, out.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        return out.cpu().numpy()


def read_stats(raws, device=None):
    """(shift, scale) = (median, MAD) of every read, float64 [R, 2]."""
    torch = _torch()
    arrays = [_as_int16(r) for r in raws]
    offsets = np.zeros(len(arrays) + 1, np.int64)
    if arrays:
        offsets[1:] = np.cumsum([a.size for a in arrays])
    raw = np.concatenate(arrays) if arrays else np.zeros(0, np.int16)
    dev = _device_index(device)
    with torch.cuda.device(dev):
        rd = torch.from_numpy(raw).to("cuda:%d" % dev)
        st = torch.empty((len(arrays), 2), dtype=torch.float64, device=rd.device)
        _cabi.check(_cabi.load_library().cf_normalize_reads(
            dev, rd.data_ptr(), _offsets_ptr(offsets), len(arrays), st.data_ptr(), None,
            torch.cuda.current_stream().cuda_stream))
        return st.cpu().numpy()


def reshape_input(data, window, n_inputs):
    """infer.py:108-124 (a reshape; no arithmetic)."""
    try:
        data = np.reshape(data, (-1, window, n_inputs))
    except ValueError:
        print(len(data))
        print(len(data[0]))
    return data


# ------------------------------------------------------------------------------ classified output
def class_from_threshold(predicted_scores, threshold=0.5):
    """infer.py:128-138: list of 0/1 labels."""
    torch = _torch()
    s = np.ascontiguousarray(np.asarray(predicted_scores, dtype=np.float64).reshape(-1))
    if s.size == 0:
        return []
    dev = _device_index()
    with torch.cuda.device(dev):
        sd = torch.from_numpy(s).to("cuda:%d" % dev)
        out = torch.empty(s.size, dtype=torch.int64, device=sd.device)
        _cabi.check(_cabi.load_library().cf_class_from_threshold(
            dev, sd.data_ptr(), s.size, float(threshold), out.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        return out.cpu().numpy().tolist()


def hp_in_pred(predictions, extension_left=11, extension_right=16, label=1):
    """infer.py:141-162: [[start - ext_left, start + len + ext_right], ...] of every run of ``label``."""
    torch = _torch()
    p = np.ascontiguousarray(np.asarray(predictions, dtype=np.int64).reshape(-1))
    if p.size == 0:
        raise IndexError("list index out of range")            # predictions[0], infer.py:151
    dev = _device_index()
    cap = p.size // 2 + 1
    with torch.cuda.device(dev):
        pd = torch.from_numpy(p).to("cuda:%d" % dev)
        out = torch.empty((cap, 2), dtype=torch.int64, device=pd.device)
        n_out = torch.zeros(1, dtype=torch.int64, device=pd.device)
        _cabi.check(_cabi.load_library().cf_hp_in_pred(
            dev, pd.data_ptr(), p.size, int(extension_left), int(extension_right), int(label),
            out.data_ptr(), cap, n_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        n = int(n_out.item())
        return out[:n].cpu().numpy().tolist()


def correct_short(predictions, threshold=15):
    """infer.py:174-198: runs of a non-zero label shorter than ``threshold`` become 0."""
    torch = _torch()
    p = np.ascontiguousarray(np.asarray(predictions, dtype=np.int64).reshape(-1))
    if p.size == 0:
        raise IndexError("list index out of range")            # predictions[0], infer.py:184
    dev = _device_index()
    with torch.cuda.device(dev):
        pd = torch.from_numpy(p).to("cuda:%d" % dev)
        out = torch.empty_like(pd)
        _cabi.check(_cabi.load_library().cf_correct_short(
            dev, pd.data_ptr(), p.size, int(threshold), out.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        return out.cpu().numpy()
