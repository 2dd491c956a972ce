"""GPU: the parity cases BASELINE.json's configs name (configs[3] is bench.py's workload,
configs[4] is tests/test_gpu_forward.py::test_ultra_long_read_vs_oracle)."""
import numpy as np
import pytest

from helpers import allowed_label_flips
from catfish_b200 import infer, neural_network, synth, weights
from oracle import postprocess, tf_graph

pytestmark = pytest.mark.gpu
PROB_TOL = 1e-3


def _build(kind, w, **hpm):
    full = dict(weights.SHIPPED_HPARAMS)
    full.update(hpm)
    m = neural_network.build_model(kind, **full)
    m.set_weights(w)
    return m


def test_config0_200_reads_of_10k_vs_oracle(shipped_weights):
    """configs[0]: ResNetRNN on 200 synthetic 10k-sample reads - every probability and every interval."""
    reads = synth.synth_reads([10000] * 200, base_seed=0)
    m = _build("ResNetRNN", shipped_weights)
    hps, lens, scores = infer.infer_reads(reads, m, return_scores=True)
    graph = tf_graph.TorchGraph(shipped_weights)
    worst, mismatched = 0.0, 0
    for r, h, n, s in zip(reads, hps, lens, scores):
        want_h, want_n, want_s = postprocess.infer_read(r, graph.infer)
        assert n == want_n == 10000
        worst = max(worst, float(np.abs(s - want_s).max()))
        if h != want_h:
            mismatched += 1
            lab_g = postprocess.class_from_threshold(s.astype(np.float64))
            lab_r = postprocess.class_from_threshold(want_s)
            diff = lab_g != lab_r
            assert diff.any() and np.all(allowed_label_flips(want_s)[diff])
            assert h == postprocess.hp_in_pred(postprocess.correct_short(lab_g))
    assert worst < PROB_TOL, worst
    assert mismatched <= 2                      # only reads with a probability inside the 1e-3 band of 0.5


def test_config1_rnn_only_20k_reads_subset_vs_oracle():
    """configs[1]: RNN-only (input width 1, H = 64 x 3), random-init, 20k-sample reads; oracle on a subset."""
    w = weights.random_init("RNN", seed=21, layer_size=64, n_layers=3)
    m = _build("RNN", w)
    reads = synth.synth_reads([20000] * 64, base_seed=2100)
    hps, lens, scores = infer.infer_reads(reads, m, return_scores=True)
    graph = tf_graph.TorchGraph(w)
    for i in (0, 17, 63):
        want_h, want_n, want_s = postprocess.infer_read(reads[i], graph.infer)
        assert lens[i] == want_n
        assert np.abs(scores[i] - want_s).max() < PROB_TOL
        if np.abs(want_s - 0.5).min() > 1e-3:
            assert hps[i] == want_h


def test_config2_resnet_only_large_batch_subset_vs_oracle():
    """configs[2]: ResNet-only window classifier at large batch (1 048 576 windows in one call, the
    tcgen05 implicit-GEMM conv path); parity on a 4 096-window subset."""
    w = weights.random_init("ResNet", seed=22, layer_size_res=32, n_layers_res=2)
    m = _build("ResNet", w)
    assert m.resolved_engine == "tcgen05"
    rng = np.random.default_rng(3)
    n_windows = 1 << 20
    x = rng.normal(0, 1.5, size=(n_windows, 35, 1)).astype(np.float32)
    p = m.infer(x)
    assert p.shape == (n_windows * 35,) and np.all(np.isfinite(p))
    sel = rng.choice(n_windows, size=4096, replace=False)
    want = tf_graph.forward_torch(w, x[sel])
    got = p.reshape(n_windows, 35)[sel].reshape(-1)
    assert np.abs(got - want).max() < PROB_TOL
    # batch invariance: the same windows alone give the same bits
    alone = m.infer(x[sel])
    np.testing.assert_array_equal(alone, got)
