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


def test_rnn_only_fused_layer0_at_scale_vs_fp32_engine():
    """RNN-only network through the fused layer kernel (scalar input folded into the epilogue, f16e5 operands in
    every layer) on a ragged batch of several thousand tiles, against the fp32 CUDA-core engine."""
    w = weights.random_init("RNN", seed=23, layer_size=64, n_layers=3)
    fast = _build("RNN", w)
    assert fast.resolved_engine == "tcgen05" and fast.operand_format == "f16e5"
    full = dict(weights.SHIPPED_HPARAMS)
    slow = neural_network.build_model("RNN", engine="simt", **full)
    slow.set_weights(w)
    reads = synth.synth_reads(synth.ragged_lengths(60, 20_000, 60_000, seed=8), base_seed=2300)
    h1, l1, s1 = infer.infer_reads(reads, fast, return_scores=True)
    h2, l2, s2 = infer.infer_reads(reads, slow, return_scores=True)
    assert l1 == l2
    worst = max(float(np.abs(a - b).max()) for a, b in zip(s1, s2))
    assert worst < PROB_TOL, worst
    for a, b, sa, sb in zip(h1, h2, s1, s2):
        if a != b:
            flip = (sa.astype(np.float64) >= 0.5) != (sb.astype(np.float64) >= 0.5)
            assert flip.any() and np.all(allowed_label_flips(sb.astype(np.float64))[flip])


def test_label_bit_path_equals_probability_path(shipped_weights):
    """With no probabilities requested the head kernel thresholds its own scores into label bits and the interval
    caller starts from those; with probabilities requested it reads them back from HBM.  Same intervals, bit for
    bit, for several thresholds, incl. a constant read (NaN scores), one-sample reads and 35-divisible lengths."""
    m = _build("ResNetRNN", shipped_weights)
    lengths = [1, 34, 35, 36, 70, 4480, 4481, 31, 64, 10000, 33333]
    reads = synth.synth_reads(lengths, base_seed=77) + [np.full(300, 5, np.int16)]
    reads += synth.synth_reads(synth.ragged_lengths(40, 3000, 90_000, seed=4), base_seed=610)
    raw, off = synth.concat_reads(reads)
    for thr, min_run, el, er in ((0.5, 15, 11, 16), (0.3, 1, 0, 0), (0.9, 40, 3, 7)):
        iv_b, ioff_b = infer.infer_concatenated(raw, off, m, threshold=thr, min_run=min_run, extension_left=el,
                                                extension_right=er)
        iv_p, ioff_p, scores = infer.infer_concatenated(raw, off, m, threshold=thr, min_run=min_run, extension_left=el,
                                                        extension_right=er, return_scores=True)
        np.testing.assert_array_equal(ioff_b, ioff_p)
        np.testing.assert_array_equal(iv_b, iv_p)
        # and both are the reference's post-processing of the returned scores
        for r in (0, 3, 9, 11, 20):
            s = scores[off[r]:off[r + 1]].astype(np.float64)
            labels = postprocess.correct_short(postprocess.class_from_threshold(s, thr), min_run)
            assert iv_b[ioff_b[r]:ioff_b[r + 1]].tolist() == postprocess.hp_in_pred(labels, el, er)


def test_sharded_job_single_rank_equals_batch_call(shipped_weights):
    """sharding.infer_reads_sharded on one rank (the N = 1 job of bench.py): small batches + CSR merge give
    exactly the per-read results of one call, through a lazy read sequence and the lazy result views."""
    from catfish_b200 import sharding
    m = _build("ResNetRNN", shipped_weights)
    lengths = synth.ragged_lengths(37, 500, 30_000, seed=12)
    reads = synth.synth_reads(lengths, base_seed=1200)
    want_h, want_l = infer.infer_reads(reads, m)

    class Lazy(object):
        def __len__(self):
            return len(reads)

        def __getitem__(self, i):
            return reads[int(i)]

    hps, lens = sharding.infer_reads_sharded(Lazy(), m, 0, 1, batch_reads=8, lengths=lengths)
    assert lens == want_l and len(hps) == len(reads)
    assert all(a == b for a, b in zip(hps, want_h))
    assert hps[5].tolist() == want_h[5].tolist() and hps[1:3][1] == want_h[2]
    # the call leaves its host wall-clock breakdown for bench.py's job leg
    assert set(sharding.last_timing) == {"partition_s", "infer_s", "gather_s"}
    assert all(v >= 0 for v in sharding.last_timing.values())


def test_reserve_gather_presizes_the_pinned_staging():
    """sharding.reserve_gather: the NCCL gather's pinned staging is sized once for the largest shard (rank 0 also
    gets the receive buffer for every rank) and later, smaller requests reuse it."""
    from catfish_b200 import sharding
    sharding._PIN.clear()
    sharding.reserve_gather(100, 5000, rank=0, world_size=4)
    need = 3 * 100 + 1 + 2 * 5000
    send, recv = sharding._PIN["send"], sharding._PIN["recv"]
    assert send.is_pinned() and send.numel() >= need and recv.numel() >= 4 * need
    sharding.reserve_gather(10, 50, rank=0, world_size=4)
    assert sharding._PIN["send"] is send and sharding._PIN["recv"] is recv
    sharding._PIN.clear()
    sharding.reserve_gather(100, 5000, rank=1, world_size=4)
    assert "recv" not in sharding._PIN
    sharding._PIN.clear()
