"""ORACLE helper (test infrastructure): import the reference's own infer.py unmodified.

/root/reference/catfish/infer.py imports h5py and the TensorFlow model classes at
module level; neither is installed.  Its pre/post-processing functions are pure
Python/numpy, so the module is imported with stub modules standing in for
``h5py``, ``models``, ``models.resnet_class`` and ``models.rnn_class``.  Nothing is
copied: the file is executed from where it lies.  Only available in the build
container (``/root/reference`` does not exist on the GPU box) - callers must check
``available()``; the tests that need it skip themselves otherwise and rely on the
committed vectors under tests/golden/ instead.
"""

import importlib.util
import os
import sys
import types

REFERENCE_INFER = "/root/reference/catfish/infer.py"
_cached = None


def available():
    return os.path.exists(REFERENCE_INFER)


def load():
    """Return the reference's infer module (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    stubs = {}
    for name, attrs in (("h5py", {}), ("models", {}),
                        ("models.resnet_class", {"ResNetRNN": object}),
                        ("models.rnn_class", {"RNN": object})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            stubs[name] = m
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_catfish_reference_infer", REFERENCE_INFER)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in stubs:
            sys.modules.pop(name, None)
    _cached = mod
    return mod


# ---------------------------------------------------------------- "next" rows N3 / N4
REFERENCE_METRICS = "/root/reference/networks/trainingDB/metrics.py"
REFERENCE_CORRECT_OUTPUT = "/root/reference/networks/correct_output.py"
_cached_more = {}


def _load_with_stubs(tag, path, stub_specs):
    if tag in _cached_more:
        return _cached_more[tag]
    if not os.path.exists(path):
        raise RuntimeError("reference tree not mounted at /root/reference")
    stubs = {}
    for name, attrs in stub_specs:
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            stubs[name] = m
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_catfish_reference_" + tag, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in stubs:
            sys.modules.pop(name, None)
    _cached_more[tag] = mod
    return mod


def load_metrics():
    """The reference's networks/trainingDB/metrics.py (plotting libraries stubbed out)."""
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    return _load_with_stubs("metrics", REFERENCE_METRICS, (
        ("matplotlib", {"use": lambda *a, **k: None}), ("matplotlib.pyplot", {}), ("seaborn", {})))


class _FakeEvents(object):
    def __init__(self, lengths):
        import numpy as np
        self._cols = {"length": np.asarray(lengths), "base": np.array(["A"] * len(lengths))}

    def __getitem__(self, key):
        return self._cols[key]


class _FakeFile(object):
    """Stands in for h5py.File: the ``read`` argument carries the event lengths."""
    def __init__(self, lengths, mode="r"):
        self._events = _FakeEvents(lengths)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __getitem__(self, path):
        return self._events


def run_correct_events(scores, event_lengths, start, length):
    """Execute the reference's correct_events (networks/correct_output.py:14-76) on in-memory
    events and recover what it computed from what it prints (it returns nothing):
    -> (classes, voted_bases, start_event, final_event).  Exceptions propagate."""
    import ast
    import contextlib
    import io
    mod = _load_with_stubs("correct_output", REFERENCE_CORRECT_OUTPUT, (
        ("base_to_signal", {"get_base_new_signal": None}), ("h5py", {"File": _FakeFile})))
    mod.h5py = types.SimpleNamespace(File=_FakeFile)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        mod.correct_events(list(event_lengths), scores, start=start, length=length)
    lines = out.getvalue().splitlines()
    classes = ast.literal_eval(next(l for l in lines if l.startswith("[")))
    tail = next(l for l in lines if l.startswith("Classified"))
    words = tail.split()
    return classes, int(words[1]), int(words[5]), int(words[7])
