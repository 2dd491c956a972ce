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
    if window_size != model.window:
        raise ValueError("window_size must equal the model's window (%d)" % model.window)
    arrays = [_as_int16(r) for r in raws]
    for a in arrays:
        if a.size == 0:
            raise IndexError("list index out of range")        # infer.py:184 on an empty read
    n_reads = len(arrays)
    offsets = np.zeros(n_reads + 1, np.int64)
    if n_reads:
        offsets[1:] = np.cumsum([a.size for a in arrays])
    total = int(offsets[-1])
    if n_reads:
        stage = _staging(_torch(), int(model.device), total).numpy()[:total]
        np.concatenate(arrays, out=stage)
        raw = stage
    else:
        raw = np.zeros(0, np.int16)
    res = infer_concatenated(raw, offsets, model, threshold, min_run, extension_left, extension_right,
                             return_scores)
    intervals, ioff = res[0], res[1]
    lengths = [int(a.size) for a in arrays]
    bounds = ioff.tolist()
    hps = [IntervalList(intervals[bounds[r]:bounds[r + 1]]) for r in range(n_reads)]
    if return_scores:
        scores = [res[2][int(offsets[r]):int(offsets[r + 1])] for r in range(n_reads)]
        return hps, lengths, scores
    return hps, lengths


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
