"""CPU: host-side logic of the package (no compute calls)."""
import os
import re

import numpy as np
import pytest

from catfish_b200 import _cabi, compute_on_read, neural_network, output_homopolymers, sharding, synth, tf_checkpoint, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CKPT = "/root/reference/catfish/ResNetRNN/checkpoints/ckpnt-30000"


def test_library_exports_every_declared_symbol():
    lib = _cabi.load_library()
    header = open(os.path.join(ROOT, "include", "catfish_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no functions parsed from the header"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cf_abi_version() == 1
    assert lib.cf_max_intervals(160, 1, 15) >= 10


def test_model_desc_validation_without_gpu():
    lib = _cabi.load_library()
    d = _cabi.ModelDesc(_cabi.NET_RESNET_RNN, 35, 64, 3, 32, 2, 1e-3, 0)
    assert lib.cf_model_num_tensors(d) == 74
    d = _cabi.ModelDesc(_cabi.NET_RNN, 35, 64, 3, 0, 0, 1e-3, 0)
    assert lib.cf_model_num_tensors(d) == 26
    d = _cabi.ModelDesc(_cabi.NET_RESNET, 35, 0, 0, 32, 2, 1e-3, 0)
    assert lib.cf_model_num_tensors(d) == 50
    bad = _cabi.ModelDesc(_cabi.NET_RNN, 34, 64, 3, 0, 0, 1e-3, 0)
    assert lib.cf_model_num_tensors(bad) == _cabi.CF_ERR_BAD_ARG
    assert b"window" in lib.cf_last_error()
    with pytest.raises(ValueError):
        _cabi.check(lib.cf_model_num_tensors(bad))


def test_weight_shapes_and_random_init():
    s = weights.expected_shapes("ResNetRNN")
    assert len(s) == 74 and sum(int(np.prod(v)) for v in s.values()) == 197185
    assert s["conv1d_2/kernel"] == (3, 32, 32) and s["conv1d/kernel"] == (1, 1, 32)
    assert s["stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/gru_cell/gates/kernel"] == (96, 128)
    assert s["stack_bidirectional_rnn/cell_1/bidirectional_rnn/bw/gru_cell/candidate/kernel"] == (192, 64)
    r = weights.expected_shapes("RNN")
    assert r["stack_bidirectional_rnn/cell_0/bidirectional_rnn/fw/gru_cell/gates/kernel"] == (65, 128)
    assert weights.expected_shapes("ResNet")["final_fully_connected/kernel"] == (32, 1)
    w = weights.random_init("ResNetRNN", seed=1)
    lim = np.sqrt(6.0 / (96 + 96))
    assert np.abs(w["conv1d_2/kernel"]).max() <= lim and np.abs(w["conv1d_2/kernel"]).max() > 0.9 * lim
    assert np.all(w["stack_bidirectional_rnn/cell_2/bidirectional_rnn/bw/gru_cell/gates/bias"] == 1.0)
    assert np.all(w["batch_normalization_3/moving_variance"] == 1.0) and np.all(w["conv1d_5/bias"] == 0.0)
    w2 = weights.random_init("ResNetRNN", seed=1)
    assert all(np.array_equal(w[k], w2[k]) for k in w)
    with pytest.raises(KeyError):
        weights.check_weights({}, "RNN")


def test_shipped_weights_match_reference_bundle(shipped_weights):
    assert len(shipped_weights) == 74
    assert np.all(shipped_weights["batch_normalization/moving_mean"] == 0)
    assert shipped_weights["final_fully_connected/kernel"].shape == (128, 1)
    if not os.path.exists(REF_CKPT + ".index"):
        pytest.skip("reference tree not mounted")
    idx = tf_checkpoint.read_index(REF_CKPT)
    assert len(idx) == 190
    w = weights.load_tf_checkpoint(REF_CKPT)          # verifies every tensor's masked crc32c
    assert set(w) == set(shipped_weights)
    for k in w:
        np.testing.assert_array_equal(w[k], shipped_weights[k])


def test_crc32c_known_answer():
    assert tf_checkpoint.crc32c(b"123456789") == 0xE3069283


def test_retrieve_hyperparams_and_build():
    hpm = neural_network.retrieve_hyperparams(os.path.join(neural_network.SHIPPED_MODEL_DIR, "ResNetRNN.txt"))
    assert hpm == weights.SHIPPED_HPARAMS
    ref_txt = "/root/reference/catfish/ResNetRNN/ResNetRNN.txt"
    if os.path.exists(ref_txt):
        assert neural_network.retrieve_hyperparams(ref_txt) == hpm
    m = neural_network.build_model("ResNetRNN", **hpm)
    assert (m.window, m.n_inputs, m.n_outputs, m.layer_sizes, m.model_type) == (35, 1, 1, [64, 64, 64], "ResNet-RNN")
    assert neural_network.build_model("RNN", **hpm).model_type == "GRU"
    assert neural_network.build_model("nonsense", **hpm) is None
    with pytest.raises(RuntimeError):
        m.infer(np.zeros((1, 35, 1)))                  # no weights restored yet


def test_partition_reads_lpt():
    lengths = synth.ragged_lengths(1000, 50000, 200000, seed=3)
    for n in (1, 2, 4, 8):
        parts = sharding.partition_reads(lengths, n)
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(1000))
        loads = np.array([lengths[p].sum() for p in parts])
        assert loads.max() - loads.min() <= lengths.max()
    assert sharding.gather_results([2, 0, 1], ["c", "a", "b"], 3, 0, 1) == ["a", "b", "c"]


def test_synth_is_deterministic():
    a, b = synth.synth_read(5000, 9), synth.synth_read(5000, 9)
    np.testing.assert_array_equal(a, b)
    assert a.dtype == np.int16 and 100 < a.min() and a.max() < 900 and 480 < np.median(a) < 520
    raw, off = synth.concat_reads(synth.synth_reads([10, 20, 30]))
    assert off.tolist() == [0, 10, 30, 60] and raw.shape == (60,)


def test_shims():
    assert compute_on_read.dict_to_ordered_list({3: "a", 1: "c", 2: "b"}) == [(1, "c"), (2, "b"), (3, "a")]
    assert compute_on_read.dict_to_ordered_list({3: "a", 1: "c", 2: "b"}, 1) == [(3, "a"), (2, "b"), (1, "c")]
    counts = {"AAAAAT": 3, "AAAAT": 2, "CCCCC": 5, "ACGTA": 1, "GGGG": 7, "TTTTTTT": 1}
    true, false = output_homopolymers.t_f_positives(counts)
    assert true == {"AAAAAT": 3, "CCCCC": 5, "TTTTTTT": 1} and false == {"AAAAT": 2, "ACGTA": 1, "GGGG": 7}
    ref = "/root/reference/networks"
    if os.path.exists(ref):
        import importlib.util
        for mod, fn, args in (("compute_on_read", "dict_to_ordered_list", ({3: "a", 1: "c"}, 1)),
                              ("output_homopolymers", "t_f_positives", (counts, 4))):
            spec = importlib.util.spec_from_file_location("_ref_" + mod, os.path.join(ref, mod + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            ours = getattr({"compute_on_read": compute_on_read, "output_homopolymers": output_homopolymers}[mod], fn)
            assert ours(*args) == getattr(m, fn)(*args)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the product package may reference it."""
    pkg = os.path.join(ROOT, "catfish_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "oracle." not in text.replace("oracle.tf_graph", "").replace("oracle/", "") or fn.endswith(".md"), fn


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device every compute entry point raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from catfish_b200 import infer
    for call in (lambda: infer.normalize_raw_signal(np.array([1, 2, 3], np.int16), "median"),
                 lambda: infer.class_from_threshold([0.1, 0.9]),
                 lambda: infer.correct_short([1, 1, 0]),
                 lambda: infer.hp_in_pred([1, 1, 0])):
        with pytest.raises(_cabi.CatfishError):
            call()
    m = neural_network.build_model("ResNetRNN", **weights.SHIPPED_HPARAMS)
    with pytest.raises(_cabi.CatfishError):
        m.set_weights(weights.load_shipped())            # cf_model_create -> CF_ERR_NO_DEVICE


def test_gather_csr_single_rank_views_and_lengths():
    """gather_csr / ShardedIntervals on one rank (no process group): per-read views in read order whatever the
    order the shard lists its reads in, lengths as a plain list, slices and iteration like a list."""
    from catfish_b200 import sharding
    idx = np.array([3, 0, 2, 1])                                   # read ids in shard order
    counts = [2, 0, 1, 3]                                          # intervals of reads 3, 0, 2, 1
    ioff = np.concatenate([[0], np.cumsum(counts)])
    flat = np.arange(2 * sum(counts), dtype=np.int64).reshape(-1, 2)
    hps, lens = sharding.gather_csr(idx, np.array([30, 0, 20, 10]), ioff, flat, 4, 0, 1)
    assert lens == [0, 10, 20, 30]
    assert [len(h) for h in hps] == [0, 3, 1, 2]
    assert hps[3].tolist() == flat[0:2].tolist() and hps[1].tolist() == flat[3:6].tolist()
    assert [h.tolist() for h in hps[1:3]] == [flat[3:6].tolist(), flat[2:3].tolist()]
    assert hps == [h.tolist() for h in hps]


def test_partition_reads_properties():
    """Every read in exactly one shard, shards sorted, deterministic, and the greedy guarantee: the heaviest shard
    is at most the mean load plus one longest read."""
    from hypothesis import given, settings, strategies as st
    from catfish_b200 import sharding

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 200_000), min_size=1, max_size=300), st.integers(1, 8))
    def check(lengths, n):
        parts = sharding.partition_reads(lengths, n)
        assert len(parts) == n
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(len(lengths)))
        assert all((np.diff(p) > 0).all() for p in parts if len(p) > 1)
        again = sharding.partition_reads(lengths, n)
        assert all((a == b).all() for a, b in zip(parts, again))
        loads = [int(np.asarray(lengths)[p].sum()) for p in parts]
        assert max(loads) <= sum(lengths) / n + max(lengths) * (1 - 1.0 / n) + 1e-9
    check()
