"""Structure-level sharding for multi-GPU inference.

Every structure is an independent graph, so the path shards by whole structures with no data-path
collective: sort by length, cut into token-budget batches of similar length (little padding), and give the
batches to ranks longest-processing-time-first.  Host-side logic only (no device work)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def batch_cost(lengths: Sequence[int], top_k: int) -> float:
    """Relative cost of one padded batch: edge work (L * min(K, L)) plus the brute-force kNN (L^2)."""
    n = max(lengths) if lengths else 0
    return float(sum(l * min(top_k, l) for l in lengths)) + 0.02 * len(lengths) * n * n


def make_batches(lengths: Sequence[int], token_budget: int) -> List[List[int]]:
    """Indices grouped into padded batches with at most ``token_budget`` padded residues each."""
    order = sorted(range(len(lengths)), key=lambda i: lengths[i], reverse=True)
    batches, cur, cur_max = [], [], 0
    for i in order:
        m = max(cur_max, lengths[i])
        if cur and m * (len(cur) + 1) > token_budget:
            batches.append(cur)
            cur, m = [], lengths[i]
        cur.append(i)
        cur_max = m
    if cur:
        batches.append(cur)
    return batches


def partition(lengths: Sequence[int], world: int, top_k: int = 30, token_budget: int = 16384) -> List[List[List[int]]]:
    """Per rank, the list of batches (lists of structure indices).  Deterministic, identical on every rank."""
    batches = make_batches(lengths, token_budget)
    costs = [batch_cost([lengths[i] for i in b], top_k) for b in batches]
    load = [0.0] * world
    out: List[List[List[int]]] = [[] for _ in range(world)]
    for j in sorted(range(len(batches)), key=lambda j: (-costs[j], j)):
        r = min(range(world), key=lambda r: (load[r], r))
        out[r].append(batches[j])
        load[r] += costs[j]
    return out


def imbalance(lengths: Sequence[int], parts: List[List[List[int]]], top_k: int = 30) -> Tuple[float, float]:
    loads = [sum(batch_cost([lengths[i] for i in b], top_k) for b in p) for p in parts]
    return max(loads), sum(loads) / len(loads)
