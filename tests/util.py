"""Shared helpers for the parity tests."""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grom_b200 import hostlib  # noqa: E402
from grom_b200.params import GA, GA_NAMES, Params  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# arrays produced by the current kernels / oracle and checked bit-exactly
PILEUP = list(range(0, 23))
CLIPS = list(range(24, 39))
DEPTH = [GA["rd_mq"], GA["rd_rd"], GA["rd_low"]]
GCS = [GA["gc"], GA["acgt"]]
EVIDENCE = list(range(0, 51))          # every per-position int of the evidence accumulation (oracle/hooks.h order)
CHECKED = EVIDENCE + DEPTH + GCS


def load_golden_fasta():
    """{lower-case contig name: uint8 chars} from tests/golden/g1.fa.gz"""
    out, name, buf = {}, None, []
    with gzip.open(os.path.join(GOLDEN, "g1.fa.gz"), "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    out[name] = np.frombuffer("".join(buf).encode(), dtype=np.uint8).copy()
                name, buf = line[1:].strip(), []
            else:
                buf.append(line.strip())
    out[name] = np.frombuffer("".join(buf).encode(), dtype=np.uint8).copy()
    return out


def golden_params(npz, rmdup):
    m = npz["mean"]
    return Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]), rmdup=rmdup)


def golden_batches():
    """(names, [ReadBatch per tid]) decoded from tests/golden/g1.bam by the host batcher."""
    with hostlib.Bam(os.path.join(GOLDEN, "g1.bam")) as b:
        return b.names, [b.read_target(t, keep_names=True) for t in range(len(b.names))]


def tables():
    """The statistics tables, computed (full doubles)."""
    return hostlib.tables(None, 20)


def tables_7digit():
    """Tables as the reference sees them when it loads its own "%e" text files (7 significant digits)."""
    hez, mq = hostlib.tables(None, 20)
    f = np.vectorize(lambda x: float("%e" % x))
    return f(hez), f(mq)


def assert_arrays_equal(got, want, names=None, where=None):
    for k in (names or CHECKED):
        a, b = got[k], want[k]
        if where is not None:
            a, b = a[where], b[where]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, f"array {GA_NAMES[k]}: {bad.size} mismatches, first at {bad[:5]}: got {a[bad[:5]]} want {b[bad[:5]]}"


def assert_clusters_equal(got, ref):
    """got = (w, rs, re, dist, mchr, other_len) from the GPU; ref = OracleResult.  read_start/read_end/dist only where a cluster exists."""
    w, rs, re, dist, mchr, ol = got
    assert np.array_equal(w, ref.cl_w), "cluster weights differ"
    live = ref.cl_w != 0
    assert np.array_equal(rs[live], ref.cl_rs[live]), "cluster read_start differs"
    assert np.array_equal(re[live], ref.cl_re[live]), "cluster read_end differs"
    assert np.array_equal(dist[live], ref.cl_dist[live]), "cluster running-mean distance differs (double, bit-exact)"
    for k in range(2):
        lv = ref.cl_w[8 + k] != 0
        assert np.array_equal(mchr[k][lv], ref.cl_mchr[k][lv]), "ctx mate contig differs"
    assert np.array_equal(ol, ref.other_len), "other_len differs"
