"""longest-processing-time greedy bin packing (mirror of distributed/assign.py:12-42): items sorted by
weight descending (in place), each put into the currently lightest bin."""
from typing import Callable, Hashable, TypeVar

__all__ = ["greedy_bin_packing"]

T = TypeVar("T", bound=Hashable)


def greedy_bin_packing(items: list, num_bins: int, item_weight_fn: Callable = lambda x: 1):
    items.sort(key=item_weight_fn, reverse=True)
    bins = [[] for _ in range(num_bins)]
    loads = [0] * num_bins
    owner = {}
    for it in items:
        b = loads.index(min(loads))
        bins[b].append(it)
        owner[it] = b
        loads[b] += item_weight_fn(it)
    return items, bins, owner
