"""Mirror of /root/reference/networks/output_homopolymers.py (``t_f_positives`` :6-32).

Host-side string analysis of k-mer counts; it is not an interval caller (the
interval logic of the hot path is in ``catfish_b200.infer``), so there is no kernel.
"""

import re


def t_f_positives(positives, k=5):
    """Split ``{kmer: count}`` into (true, false) homopolymer dicts: a k-mer is a true
    positive when it is at least ``k`` long and starts with ``k`` identical bases."""
    starts_with_run = re.compile("A{%d}|C{%d}|G{%d}|T{%d}" % (k, k, k, k))
    true, false = {}, {}
    for kmer, count in positives.items():
        if len(kmer) >= k and starts_with_run.match(kmer):
            true[kmer] = count
        else:
            false[kmer] = count
    print("TPs: ", sum(true.values()))
    print("FPs: ", sum(false.values()))
    return true, false
