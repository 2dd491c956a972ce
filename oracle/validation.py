"""ORACLE (test infrastructure, never on the product path): validation counting and event voting.

CPU restatement of
  * RNN.test_network, networks/rnn_class.py:222-261, with compute_loss :72-77 and
    compute_accuracy :80-86 (TensorFlow ops restated in numpy float64) and
    metrics.confusion_matrix, networks/trainingDB/metrics.py:10-37;
  * the loop of correct_events, networks/correct_output.py:38-61.
Pinning: ``confusion_counts`` and ``vote_events`` are checked against the reference's own
functions executed in the build container (oracle/ref_infer.py ``load_metrics`` /
``run_correct_events``; vectors committed under tests/golden/).  The loss / accuracy formulas
are pinned by the reference's shipped meta-graph: its ``accuracy/Mean`` and ``loss/Mean`` nodes
executed by tests/tools/meta_graph_interp.py (golden ``val_*`` entries of
tests/golden/forward_resnetrnn_shipped.npz; tests/test_oracle.py::test_validation_heads_against_golden).
"""

import numpy as np


def confusion_counts(true_labels, predicted_labels):
    """metrics.py:10-37 as a plain scan."""
    if len(true_labels) != len(predicted_labels):
        raise ValueError("Length of labels to compare is not equal.")
    tp = fp = tn = fn = 0
    for truth, call in zip(true_labels, predicted_labels):
        if call == 1:
            tp, fp = (tp + 1, fp) if truth == 1 else (tp, fp + 1)
        elif call == 0:
            tn, fn = (tn + 1, fn) if truth == 0 else (tn, fn + 1)
    return tp, fp, tn, fn


def test_network(logits, labels, padding_size, threshold=0.5):
    """-> ((tp, fp, tn - padding, fn), accuracy, loss) for one read (rnn_class.py:222-261).

    ``logits`` are the dense layer's outputs for every position of the padded windows,
    ``labels`` the padded labels, both flattened."""
    z = np.asarray(logits, dtype=np.float64).reshape(-1)
    y = np.asarray(labels, dtype=np.float64).reshape(-1)
    p32 = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)      # the graph is fp32 (:84)
    conf = p32.astype(float)                                  # .astype(float), :232
    pred = (conf >= threshold).astype(np.int64)               # :235
    tp = int(np.count_nonzero((pred == 1) & (y == 1)))
    fp = int(np.count_nonzero(pred == 1)) - tp
    tn = int(np.count_nonzero((pred == 0) & (y == 0)))
    fn = int(np.count_nonzero(pred == 0)) - tn
    accuracy = float(np.mean(np.rint(p32) == y))              # tf.round: half to even
    loss = float(np.mean(np.maximum(z, 0.0) - z * y + np.log1p(np.exp(-np.abs(z)))))
    return (tp, fp, tn - int(padding_size), fn), accuracy, loss


def vote_events(scores, event_lengths, start, length):
    """correct_output.py:38-61 -> (classes, start_event, final_event); unbound names -> None."""
    scores = [float(s) for s in scores]
    position = 0
    first = last = None
    span_begin = None
    classes = []
    for n, ev_len in enumerate(event_lengths):
        if first is None and position >= start:
            first, span_begin = n, position
        position += int(ev_len)
        if position > length:
            last = n - 1
            break
        if first is not None:
            members = scores[span_begin:position]
            classes.append(round(sum(members) / len(members)))
            span_begin = position
            if position == start + length:
                last = n
                break
    return classes, first, last
