"""Validation counts (next row N4) and event voting (next row N3).

CPU: the oracle restatements against vectors produced by the reference's own functions
(tests/tools/make_golden_next.py) and the host-side helpers.  GPU: the K9 / K10 kernels through the
C ABI against the oracle and the golden vectors."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, allowed_label_flips
from oracle import validation


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def _vote_or_exception(fn, c):
    try:
        return fn(c["scores"], c["events"], c["start"], c["length"]), None
    except ZeroDivisionError:
        return None, "ZeroDivisionError"
    except UnboundLocalError:
        return None, "UnboundLocalError"


def _check_vote(res, exc, c):
    if c["raises"] == "ZeroDivisionError":
        assert exc == "ZeroDivisionError"
    elif c["raises"] == "UnboundLocalError":
        assert exc == "UnboundLocalError" or (res is not None and (res[1] is None or res[2] is None))
    else:
        assert exc is None
        assert res[0] == c["classes"] and len(res[0]) == c["voted"]
        assert res[1] == c["start_event"] and res[2] == c["final_event"]


def test_oracle_vote_matches_reference_vectors():
    cases = _load("events.json")
    assert sum(c["raises"] is None for c in cases) > 50
    for c in cases:
        res, exc = _vote_or_exception(validation.vote_events, c)
        _check_vote(res, exc, c)


def test_confusion_oracle_and_host_helper_match_reference_vectors():
    from catfish_b200 import metrics
    for c in _load("confusion.json"):
        assert list(validation.confusion_counts(c["true"], c["pred"])) == c["counts"]
        assert list(metrics.confusion_matrix(c["true"], c["pred"])) == c["counts"]
    with pytest.raises(ValueError):
        metrics.confusion_matrix([1, 0], [1])


def test_metric_scalars():
    from catfish_b200 import metrics
    assert metrics.precision_recall(3, 1, 2) == (0.75, 0.6)
    assert metrics.precision_recall(0, 0, 0) == (0, 0)
    assert metrics.calculate_accuracy(1, 1, 1, 1) == 0.5 and metrics.calculate_accuracy(0, 0, 0, 0) == 0
    assert metrics.f1(0.5, 0.5) == 0.5 and metrics.f1(0, 0) == 0
    assert metrics.weighted_f1(0.5, 0.5, 1, 4) == 0.125


def test_validation_padding_rule():
    from catfish_b200 import train_validate
    x, pad = train_validate.padding(np.arange(70))
    assert pad == 0 and x.shape == (2, 35, 1)              # no extra window, unlike infer.py:32-38
    x, pad = train_validate.padding(np.arange(71))
    assert pad == 34 and x.shape == (3, 35, 1) and x[2, 1:, 0].sum() == 0


def test_oracle_test_network_hand_case():
    z = np.array([3.0, -2.0, 0.0, 1.0, -4.0])
    y = np.array([1, 1, 0, 0, 0])
    counts, acc, loss = validation.test_network(z, y, padding_size=1)
    # p = .95 .12 .5 .73 .02 -> pred (>= 0.5) 1 0 1 1 0 ; round-half-even(0.5) = 0
    assert counts == (1, 2, 1 - 1, 1)
    assert acc == pytest.approx(3 / 5)
    want = np.mean([np.log1p(np.exp(-3)), 2 + np.log1p(np.exp(-2)), np.log(2), 1 + np.log1p(np.exp(-1)), np.log1p(np.exp(-4))])
    assert loss == pytest.approx(want, rel=1e-12)


# ------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_vote_matches_reference_vectors():
    from catfish_b200 import correct_output
    for c in _load("events.json"):
        res, exc = _vote_or_exception(correct_output.vote_events, c)
        if c["raises"] == "UnboundLocalError":
            assert exc == "UnboundLocalError"
        else:
            _check_vote(res, exc, c)


@pytest.mark.gpu
def test_gpu_vote_long_read_vs_oracle():
    from catfish_b200 import correct_output
    rng = np.random.default_rng(5)
    ev = rng.integers(1, 30, 60000)
    total = int(ev.sum())
    scores = rng.random(total).astype(np.float32).astype(np.float64)
    cum = np.cumsum(ev)
    start, length = int(cum[99]), int(cum[-200])
    got = correct_output.vote_events(scores, ev, start, length)
    want = validation.vote_events(scores, ev, start, length)
    assert got[0] == want[0] and got[1:] == want[1:]
    assert len(got[0]) == 60000 - 199 - 100


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["simt", "auto"])
def test_gpu_test_network_vs_oracle(engine, shipped_weights):
    from catfish_b200 import neural_network, train_validate, weights
    from oracle import tf_graph
    rng = np.random.default_rng(11)
    m = neural_network.build_model("ResNetRNN", engine=engine, **weights.SHIPPED_HPARAMS)
    m.set_weights(shipped_weights)
    tot = np.zeros(4, np.int64)
    for n in (35 * 4, 2000, 4513):
        data = rng.normal(0, 1.5, n)
        labels = (rng.random(n) < 0.3).astype(np.int64)
        x, pad = train_validate.padding(data)
        y, _ = train_validate.padding(labels)
        before = np.array([m.tp, m.fp, m.tn, m.fn])
        acc, loss = m.test_network(x, y, "read", None, pad)
        counts = np.array([m.tp, m.fp, m.tn, m.fn]) - before
        z = tf_graph.forward_np(shipped_weights, x.astype(np.float32), return_logits=True).reshape(-1)
        (tp, fp, tn, fn), acc_ref, loss_ref = validation.test_network(z, y, pad)
        p_ref = 1.0 / (1.0 + np.exp(-z))
        slack = int(allowed_label_flips(p_ref).sum())
        assert np.abs(counts - np.array([tp, fp, tn, fn])).max() <= slack
        assert counts.sum() == x.shape[0] * 35 - pad
        assert abs(float(acc) - acc_ref) <= (slack + 0.5) / z.size + 1e-6
        assert abs(float(loss) - loss_ref) < 2e-4 * max(1.0, loss_ref)
        # the counts are exactly the counting of the library's own probabilities
        probs = m.infer(x)
        pred = (probs >= 0.5).astype(np.int64)
        own = validation.confusion_counts(y.reshape(-1).tolist(), pred.tolist())
        assert list(counts) == [own[0], own[1], own[2] - pad, own[3]]
        assert float(acc) == pytest.approx(np.mean(np.rint(probs.astype(np.float32)) == y.reshape(-1)), abs=1e-7)
        tot += counts
    assert [m.tp, m.fp, m.tn, m.fn] == list(tot)


@pytest.mark.gpu
def test_gpu_validate_driver(shipped_weights):
    from catfish_b200 import neural_network, train_validate, weights
    rng = np.random.default_rng(3)
    m = neural_network.build_model("ResNetRNN", **weights.SHIPPED_HPARAMS)
    m.set_weights(shipped_weights)
    reads = [(rng.normal(0, 1.5, n), (rng.random(n) < 0.2).astype(int)) for n in (900, 1500, 300)]
    acc, prec, rec = train_validate.validate(m, reads, 700, None, validation_start=0)
    assert 0.0 <= acc <= 1.0 and 0.0 <= prec <= 1.0 and 0.0 <= rec <= 1.0
    assert (m.tp, m.fp, m.tn, m.fn) == (0, 0, 0, 0)          # counters reset (:283-286)
    # the 300-sample read is skipped (< start + max length); 2 reads x 700 positions were counted
    m2 = neural_network.build_model("ResNetRNN", **weights.SHIPPED_HPARAMS)
    m2.set_weights(shipped_weights)
    for data, labels in reads[:2]:
        x, pad = train_validate.padding(data[:700])
        y, _ = train_validate.padding(labels[:700])
        m2.test_network(x, y, None, None, pad)
    assert m2.tp + m2.fp + m2.tn + m2.fn == 1400
    from catfish_b200 import metrics
    assert acc == metrics.calculate_accuracy(m2.tp, m2.fp, m2.tn, m2.fn)
