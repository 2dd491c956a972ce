"""catfish_b200 - B200-native inference hot path of MMAThijssen/catfish.

Drop-in for the reference's per-read entry points: the modules mirror the
reference's names (``infer``, ``neural_network``, ``rnn_class``, ``resnet_class``,
``compute_on_read``, ``output_homopolymers``); all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/catfish_b200.h``.
PyTorch tensors only carry device buffers and streams.  There is no CPU fallback.
"""

__version__ = "0.1.0"

_default_device = 0


def set_device(index):
    """CUDA device used by the module-level helpers (models carry their own)."""
    global _default_device
    _default_device = int(index)


def get_device():
    return _default_device
