"""Host-side helpers of bench.py that decide without a GPU: the cores every rank of a node binds to (bind_rank_to_cores -> plan_cores)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _cpulist, bind_rank_to_cores, plan_cores  # noqa: E402


def test_cpulist():
    assert _cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and _cpulist("") == set()


def test_ranks_get_disjoint_cores_on_the_node_of_their_gpu():
    nodes = [0, 0, 0, 0, 1, 1, 1, 1]
    node_cpus = {0: set(range(0, 16)) | set(range(32, 48)), 1: set(range(16, 32)) | set(range(48, 64))}
    got = [plan_cores(range(64), nodes, node_cpus, i) for i in range(8)]
    assert sorted(c for g in got for c in g[0]) == list(range(64))                       # disjoint, nothing left idle
    for i, (cores, node, sharing) in enumerate(got):
        assert node == nodes[i] and sharing == 4 and set(cores) <= node_cpus[node] and len(cores) == 8
    # the container may use one socket only: every rank still gets its own cores
    got = [plan_cores(range(16), nodes, node_cpus, i)[0] for i in range(8)]
    assert sorted(c for g in got for c in g) == list(range(16)) and all(len(g) == 2 for g in got)
    # unknown topology: an even split of what is allowed
    got = [plan_cores([3, 4, 5, 9, 10, 11], [-1, -1, -1], {}, i)[0] for i in range(3)]
    assert got == [[3, 4], [5, 9], [10, 11]]
    # two ranks, one GPU node known and one not: all ranks must decide alike (split everything)
    got = [plan_cores(range(8), [0, -1], {0: set(range(8))}, i)[0] for i in range(2)]
    assert got == [[0, 1, 2, 3], [4, 5, 6, 7]]


def test_binding_never_fails_the_run_without_a_gpu():
    before = os.sched_getaffinity(0)
    assert bind_rank_to_cores(0, 1) is None
    r = bind_rank_to_cores(1, 2)                 # no CUDA device here: reported, nothing changed
    assert (r is None or "error" in r) and os.sched_getaffinity(0) == before
