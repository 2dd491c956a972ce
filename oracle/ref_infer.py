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
