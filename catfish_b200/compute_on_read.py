"""Mirror of /root/reference/networks/compute_on_read.py.

Only ``dict_to_ordered_list`` (:3-18) is computable in the reference; its second
function, ``average_between_predictions`` (:21-45), uses undefined names and
cannot run, so it is not mirrored.  Host-side list handling, no kernel.
"""


def dict_to_ordered_list(dict_in, sort_on=0):
    """List of (key, value) tuples sorted on element ``sort_on`` (0 = key, 1 = value)."""
    return sorted(dict_in.items(), key=lambda item: item[sort_on])
