"""Per-event majority vote ("next" row N3).

Reference: /root/reference/networks/correct_output.py ``correct_events`` :14-76, which the reference
marks "NOT FINISHED": it computes, for the Tombo events between ``start`` and ``length``, the class
``round(mean(scores of the event's measurements))`` and then stops without returning it.
``vote_events`` is the array-level twin (event lengths in, classes out; the vote runs on the GPU
behind ``cf_vote_events``); ``correct_events`` reads the events from a resquiggled FAST5 with h5py
and returns what the reference computed.
"""

import ctypes

import numpy as np

from . import _cabi


def vote_events(scores, event_lengths, start=30000, length=4970, device=None):
    """-> (classes list[int], start_event, final_event) as in correct_output.py:38-61.

    Raises what the reference's loop raises: ``UnboundLocalError`` when it would print an unbound
    ``start_event`` / ``final_event`` (:68), ``ZeroDivisionError`` for an event without scores (:57).
    """
    import torch
    from . import get_device
    dev = torch.device("cuda", get_device() if device is None else int(device))
    sc = np.ascontiguousarray(np.asarray(scores, dtype=np.float64))
    ev = np.ascontiguousarray(np.asarray(event_lengths, dtype=np.int64))
    n_voted, first, final = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    empty = ctypes.c_int32()
    with torch.cuda.device(dev):
        sd = torch.from_numpy(sc).to(dev)
        cd = torch.empty(max(1, ev.size), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev)
        _cabi.check(_cabi.load_library().cf_vote_events(
            dev.index, sd.data_ptr(), sc.size, ev.ctypes.data_as(_cabi.c_i64_p), ev.size, int(start), int(length),
            cd.data_ptr(), ctypes.byref(n_voted), ctypes.byref(first), ctypes.byref(final), ctypes.byref(empty),
            stream.cuda_stream))
        classes = cd[: n_voted.value].cpu().numpy().tolist()
    if empty.value:
        raise ZeroDivisionError("division by zero")
    if first.value < 0:
        raise UnboundLocalError("local variable 'start_event' referenced before assignment")
    if final.value == -2:
        raise UnboundLocalError("local variable 'final_event' referenced before assignment")
    return classes, first.value, final.value


def correct_events(read, scores, start=30000, length=4970, use_tombo=True):
    """correct_output.py:14-76 with the classes returned (the reference returns nothing)."""
    import h5py
    if not use_tombo:
        raise Exception("Not yet implemented for uncorrected reads")
    with h5py.File(read, "r") as hdf:
        events = hdf["Analyses/RawGenomeCorrected_000/BaseCalled_template/Events"]
        event_lengths = np.asarray(events["length"])
    classes, start_event, final_event = vote_events(scores, event_lengths, start, length)
    if len(classes) != len(event_lengths[start_event: final_event + 1]):
        raise ValueError("Number of events is not equal to number of classified events")
    return classes
