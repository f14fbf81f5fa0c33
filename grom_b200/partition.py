"""Chromosome -> rank assignment, mirroring the reference's `-P` policy.

The reference sorts the FASTA contigs by length, largest first (reference src/GROM.c:22318-22336), launches one
child process per contig as slots become free (src/GROM.c:549-599) and concatenates the per-contig outputs in BAM
header order (src/GROM.c:21121-21126).  Chromosomes never exchange data, so the multi-GPU form is
the same policy with "slot" = GPU: largest-first greedy assignment to the least-loaded rank, no data-path collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence


def skip_contig(name: str, gender: int) -> bool:
    """chrY / y is skipped when -g 0 (reference src/GROM.c:20979-20988); names are compared lower-cased."""
    n = name.lower()
    return gender == 0 and n in ("chry", "y")


def assign_contigs(lengths: Sequence[int], world_size: int, weights: Optional[Sequence[float]] = None) -> List[List[int]]:
    """Largest-first greedy (LPT) assignment.  Returns, per rank, the contig indices it owns in the order it should
    process them (largest first).  `weights` (e.g. read counts from the BAI metadata bin) override lengths as load."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    load = list(weights) if weights is not None else [float(x) for x in lengths]
    order = sorted(range(len(lengths)), key=lambda i: (-load[i], i))
    ranks: List[List[int]] = [[] for _ in range(world_size)]
    tot = [0.0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: (tot[k], k))
        ranks[r].append(i)
        tot[r] += load[i]
    return ranks


def merge_in_header_order(per_rank: Sequence[Dict[int, str]]) -> str:
    """Concatenate per-contig output text in BAM header (tid) order, as the parent process does with `cat >>`."""
    merged: Dict[int, str] = {}
    for d in per_rank:
        for tid, text in d.items():
            if tid in merged:
                raise ValueError(f"contig {tid} produced by two ranks")
            merged[tid] = text
    return "".join(merged[t] for t in sorted(merged))
