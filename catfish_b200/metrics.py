"""Host-side scalar metrics of the validation path ("next" row N4).

Reference: /root/reference/networks/trainingDB/metrics.py (``confusion_matrix`` :10-37,
``precision_recall`` :40-55, ``calculate_accuracy`` :58-63, ``weighted_f1`` :95-113, ``f1`` :117-135).
The per-position counting for a forward pass runs on the GPU inside ``RNN.test_network``
(C ABI ``cf_validate_windows``); ``confusion_matrix`` here is the array-level twin for label
lists that already live on the host.  Plotting helpers are out of scope.
"""

import numpy as np


def confusion_matrix(true_labels, predicted_labels):
    """(true_pos, false_pos, true_neg, false_neg); predictions other than 0 / 1 are not counted."""
    if len(true_labels) != len(predicted_labels):
        raise ValueError("Length of labels to compare is not equal.")
    t = np.asarray(true_labels)
    p = np.asarray(predicted_labels)
    called, uncalled = p == 1, p == 0
    true_pos = int(np.count_nonzero(called & (t == 1)))
    false_pos = int(np.count_nonzero(called)) - true_pos
    true_neg = int(np.count_nonzero(uncalled & (t == 0)))
    false_neg = int(np.count_nonzero(uncalled)) - true_neg
    return true_pos, false_pos, true_neg, false_neg


def _ratio(num, den, complaint=None):
    if den == 0:
        if complaint:
            print(complaint)
        return 0
    return num / den


def precision_recall(true_pos, false_pos, false_neg):
    precision = _ratio(true_pos, true_pos + false_pos, "Precision could not be calculated.")
    recall = _ratio(true_pos, true_pos + false_neg, "Recall could not be calculated.")
    return precision, recall


def calculate_accuracy(true_pos, false_pos, true_neg, false_neg):
    return _ratio(true_pos + true_neg, true_pos + false_pos + true_neg + false_neg)


def f1(precision, recall):
    return _ratio(2 * (precision * recall), precision + recall,
                  "Precision, recall or both are zero. Unable of calculating weighted F1.")


def weighted_f1(precision, recall, n, N):
    return _ratio(2 * n / N * (precision * recall), precision + recall,
                  "Precision, recall or both are zero. Unable of calculating weighted F1.")
