"""CPU: pin the oracle against the reference's own functions and the golden vectors."""
import numpy as np
import pytest

import os
import sys

from helpers import FakeFast5, golden, hpm_from_golden, random_bn_weights, random_labels
from catfish_b200 import synth, weights
from oracle import postprocess, ref_infer, tf_graph, validation

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
import meta_graph_interp as mgi  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_infer.available(), reason="reference tree not mounted")
needs_meta = pytest.mark.skipif(not mgi.available(), reason="reference meta-graph not mounted")


def test_postprocess_against_golden():
    g = golden("postprocess.npz")
    for i in range(int(g["n_patterns"])):
        p = g["pat%d" % i]
        for fn in (postprocess.correct_short, postprocess.correct_short_loops):
            np.testing.assert_array_equal(fn(list(p)), g["pat%d_correct_short" % i])
        for lab in (1, 0):
            want = g["pat%d_hp_label%d" % (i, lab)].tolist()
            assert postprocess.hp_in_pred(p, label=lab) == want
            assert postprocess.hp_in_pred_loops(list(p), label=lab) == want
        assert postprocess.hp_in_pred(p, 3, 0) == g["pat%d_hp_ext" % i].tolist()
    s = g["scores"]
    np.testing.assert_array_equal(postprocess.class_from_threshold(s), g["scores_labels_0.5"])
    np.testing.assert_array_equal(postprocess.class_from_threshold(s, 0.9), g["scores_labels_0.9"])
    assert postprocess.class_from_threshold_loops(s) == g["scores_labels_0.5"].tolist()
    for i in range(int(g["n_raws"])):
        with np.errstate(all="ignore"):
            got = postprocess.normalize_raw_signal(g["raw%d" % i])
        np.testing.assert_array_equal(got, g["raw%d_norm" % i])      # bit-exact, NaN == NaN


def test_known_answers_from_survey():
    # SURVEY.md section 8c, observed on the reference's functions
    assert postprocess.hp_in_pred([1] * 20 + [0] * 5) == [[-11, 36]]
    assert postprocess.hp_in_pred([0] * 5 + [1] * 20) == [[-6, 41]]
    out = postprocess.correct_short([1] * 14 + [0] * 3 + [1] * 15)
    assert out.tolist() == [0] * 17 + [1] * 15
    np.testing.assert_array_equal(postprocess.normalize_raw_signal(np.array([1, 2, 3, 4])), [-1.5, -0.5, 0.5, 1.5])
    assert [postprocess.padding_size(n) for n in (1, 34, 35, 36, 70, 10000)] == [34, 1, 35, 34, 35, 10]
    with pytest.raises(IndexError):
        postprocess.hp_in_pred([])


@needs_ref
def test_postprocess_against_reference_functions():
    ref = ref_infer.load()
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 3000))
        labels = random_labels(rng, n, p_switch=float(rng.choice([0.02, 0.08, 0.3])))
        np.testing.assert_array_equal(postprocess.correct_short(labels), ref.correct_short(list(labels)))
        assert postprocess.hp_in_pred(labels) == ref.hp_in_pred(list(labels))
        thr = int(rng.integers(1, 40))
        np.testing.assert_array_equal(postprocess.correct_short(labels, thr), ref.correct_short(list(labels), thr))
        scores = rng.random(n)
        assert postprocess.class_from_threshold(scores).tolist() == ref.class_from_threshold(scores)
        raw = synth.synth_read(n, 900 + trial)
        with np.errstate(all="ignore"):
            np.testing.assert_array_equal(postprocess.normalize_raw_signal(raw), ref.normalize_raw_signal(raw, "median"))


@needs_ref
def test_infer_read_matches_reference_driver_logic(shipped_weights):
    """The oracle's infer_read == the statements of infer_class_from_signal (infer.py:30-51)
    executed with the reference's own helper functions around the same network callable."""
    ref = ref_infer.load()
    g = tf_graph.TorchGraph(shipped_weights)
    for n in (700, 1225):
        raw = synth.synth_read(n, 31 + n)
        norm = ref.normalize_raw_signal(raw, "median")
        if not (len(norm) / 35).is_integer():
            pad = 35 - (len(norm) - (len(norm) // 35 * 35))
        else:
            pad = 35
        x = ref.reshape_input(np.hstack((norm, np.array(pad * [0]))), 35, 1)
        scores = g.infer(x)[:-pad]
        labels = ref.correct_short(ref.class_from_threshold(scores))
        want = ref.hp_in_pred(labels)
        hps, length, got_scores = postprocess.infer_read(raw, g.infer)
        assert hps == want and length == len(labels) == n
        np.testing.assert_array_equal(got_scores, scores)
        hps_loops, _, _ = postprocess.infer_read(raw, g.infer, loops=True)
        assert hps_loops == want


def test_forward_against_golden_shipped(shipped_weights):
    g = golden("forward_resnetrnn_shipped.npz")
    graph = tf_graph.TorchGraph(shipped_weights)
    for i in range(int(g["n_reads"])):
        raw = g["read%d" % i]
        hps, length, scores = postprocess.infer_read(raw, graph.infer)
        assert length == len(raw)
        assert np.abs(scores - g["read%d_p64" % i]).max() < 2e-6       # fp32 graph vs fp64 graph
        assert np.abs(scores - g["read%d_p32" % i]).max() < 2e-6       # thread-count dependent summation
        # intervals: identical unless a probability sits within 1e-6 of the threshold
        if np.abs(g["read%d_p64" % i] - 0.5).min() > 1e-5:
            assert hps == g["read%d_hps" % i].tolist()
        x, pad = postprocess.pad_and_window(postprocess.normalize_raw_signal(raw))
        p64 = tf_graph.forward_np(shipped_weights, x, np.float64)[:-pad]
        np.testing.assert_allclose(p64, g["read%d_p64" % i], rtol=0, atol=1e-12)


def test_golden_sources_are_the_reference_graph():
    """The probability goldens come from executing the reference's shipped meta-graph, not from the oracle."""
    for name in ("forward_resnetrnn_shipped.npz", "forward_resnetrnn_randbn_seed14.npz",
                 "forward_rnn_le64_ns3_seed11.npz", "forward_resnet_ls32_ns2_seed12.npz"):
        assert str(golden(name)["source"]).startswith("meta_graph_interp")


def test_forward_against_golden_random_bn():
    """Seeded weights with non-trivial BN statistics through the reference graph (golden) vs the oracle."""
    g = golden("forward_resnetrnn_randbn_seed14.npz")
    w = random_bn_weights(int(g["seed"]))
    np.testing.assert_allclose(tf_graph.forward_np(w, g["x"], np.float64), g["p64"], rtol=0, atol=1e-12)
    assert np.abs(tf_graph.forward_torch(w, g["x"]) - g["p64"]).max() < 2e-6
    assert np.abs(g["p32"] - g["p64"]).max() < 2e-6


def test_validation_heads_against_golden(shipped_weights):
    """accuracy/Mean and loss/Mean of the reference graph (rnn_class.py:73-88, 240-241) vs oracle.validation."""
    g = golden("forward_resnetrnn_shipped.npz")
    x, pad = postprocess.pad_and_window(postprocess.normalize_raw_signal(g["read0"]))
    z = tf_graph.forward_np(shipped_weights, x, np.float64, return_logits=True)
    _, acc, loss = validation.test_network(z, g["val_labels"], pad)
    assert acc == float(g["val_acc64"])
    assert abs(loss - float(g["val_loss64"])) < 1e-12
    assert abs(acc - float(g["val_acc32"])) < 1e-6 and abs(loss - float(g["val_loss32"])) < 1e-6


@needs_meta
def test_oracle_equals_reference_meta_graph_live():
    """Pin: oracle.tf_graph == the reference's ckpnt-30000.meta executed op by op, on 3 reads incl. a
    35-divisible length (the extra-window padding case), fp64 to 1e-12 and torch-fp32 to 2e-6."""
    ref_vars = mgi.load_reference_variables()
    shipped = weights.load_shipped()
    assert sorted(k for k in ref_vars if weights_is_inference(k)) == sorted(shipped)
    tg = tf_graph.TorchGraph(shipped)
    graph64 = mgi.MetaGraph(mgi.META, ref_vars, np.float64)
    assert graph64.tf_version == "1.10.0" and len(graph64.subgraph(mgi.PREDICTIONS)) == 1049
    for n, seed in ((350, 1), (613, 2), (1999, 3)):
        raw = synth.synth_read(n, seed)
        x, pad = postprocess.pad_and_window(postprocess.normalize_raw_signal(raw))
        p64, _ = mgi.predictions(x, graph=graph64)
        assert set(graph64.iterations.values()) == {35} and len(graph64.iterations) == 6
        np.testing.assert_allclose(tf_graph.forward_np(shipped, x, np.float64), p64, rtol=0, atol=1e-12)
        assert np.abs(tg.infer(x) - p64).max() < 2e-6
    p32, g32 = mgi.predictions(x, np.float32, variables=ref_vars, seed=5)
    assert np.abs(p32 - p64).max() < 2e-6
    assert len(g32.executed) == 44                                      # every op type of the sub-graph ran
    # dropout mask at keep_prob 1.0 is the identity whatever the random draw (rnn_class.py:151-154, 217)
    np.testing.assert_array_equal(mgi.predictions(x, np.float32, variables=ref_vars, seed=6)[0], p32)


@needs_meta
@pytest.mark.parametrize("kind,hpm", [("RNN", dict(layer_size=64, n_layers=3)),
                                      ("ResNet", dict(layer_size_res=32, n_layers_res=2))])
def test_variants_equal_rewired_reference_sub_graphs(kind, hpm):
    w = weights.random_init(kind, seed=21, **hpm)
    x = np.random.default_rng(3).normal(0, 1.5, size=(6, 35, 1)).astype(np.float32)
    np.testing.assert_allclose(tf_graph.forward_np(w, x, np.float64), mgi.predictions_variant(kind, x, w),
                               rtol=0, atol=1e-12)


def weights_is_inference(name):
    from catfish_b200 import tf_checkpoint
    return tf_checkpoint.is_inference_tensor(name)


@pytest.mark.parametrize("name", ["forward_rnn_le64_ns3_seed11.npz", "forward_resnet_ls32_ns2_seed12.npz",
                                  "forward_rnn_le16_ns2_seed13.npz"])
def test_forward_against_golden_random_init(name):
    g = golden(name)
    w = weights.random_init(str(g["kind"]), seed=int(g["seed"]), **hpm_from_golden(g))
    p64 = tf_graph.forward_np(w, g["x"], np.float64)
    np.testing.assert_allclose(p64, g["p64"], rtol=0, atol=1e-12)
    p32 = tf_graph.forward_torch(w, g["x"])
    assert np.abs(p32 - g["p64"]).max() < 2e-6
    assert tf_graph.describe(w)[0] == str(g["kind"])


def test_gru_is_reset_before_matmul():
    """TF GRUCell applies r before the candidate matmul; torch.nn.GRU applies it after."""
    rng = np.random.default_rng(0)
    w = weights.random_init("RNN", seed=3, layer_size=8, n_layers=1)
    x = rng.normal(size=(2, 35, 1)).astype(np.float32)
    p = tf_graph.forward_np(w, x, np.float64)
    # hand-rolled single step for window 0, t = 0, forward direction
    pre = "stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/gru_cell"
    wg, bg = w[pre + "/gates/kernel"].astype(float), w[pre + "/gates/bias"].astype(float)
    wc, bc = w[pre + "/candidate/kernel"].astype(float), w[pre + "/candidate/bias"].astype(float)
    h = np.zeros(8)
    xs = x[0, 0].astype(float)
    g_ = 1 / (1 + np.exp(-(np.concatenate([xs, h]) @ wg + bg)))
    r, u = g_[:8], g_[8:]
    c = np.tanh(np.concatenate([xs, r * h]) @ wc + bc)
    h1 = u * h + (1 - u) * c
    out = tf_graph._gru_direction_np(x.astype(float), w, pre, False)
    np.testing.assert_allclose(out[0, 0], h1, atol=1e-12)
    assert p.shape == (70,)


@needs_ref
def test_fast5_trimming_matches_reference_process_signal():
    """process_signal (infer.py:77-93) run by the REFERENCE's own function on an in-memory FAST5 stand-in equals
    the product's host-side trimming (`_trimmed_raw`: first_sample_template, first Raw/Reads member) followed by
    the oracle's normalisation."""
    from catfish_b200 import infer as product_infer
    ref = ref_infer.load()
    for n, first in ((900, 0), (900, 137), (36, 1)):
        f = FakeFast5(synth.synth_read(n, 40 + n + first), first, read_names=("Read_5", "Read_9"))
        want = ref.process_signal(f, "median")
        trimmed = product_infer._trimmed_raw(f)
        assert trimmed.dtype == np.int16 and len(trimmed) == n - first
        np.testing.assert_array_equal(trimmed, f.signal[first:])
        np.testing.assert_array_equal(postprocess.normalize_raw_signal(trimmed), want)
