"""Pins the CNV oracle (oracle/grom_oracle_cnv.c) against the reference's own outputs stored in tests/golden/g2_cnv.npz
(made by tests/golden/make_golden.py from the white-box build's dump hooks and checked there against the prebuilt binary)."""
import os

import numpy as np

from util import GOLDEN
from grom_b200 import hostlib
from grom_b200.params import CNV_CALL_DTYPE, Params
from oracle import pyoracle as po


def load():
    g = np.load(os.path.join(GOLDEN, "g2_cnv.npz"))
    m = g["mean"]
    prm = Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]), rand_seed=int(g["seed"]))
    return g, prm


def run_oracle(g, prm):
    return po.cnv_run(prm, "chra", g["fasta"], g["gc"].astype(np.int32), g["acgt"].astype(np.int32), g["rd_mq"].astype(np.int32),
                      g["rd_rd"].astype(np.int32), g["rd_low"].astype(np.int32), seed=prm.rand_seed)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | ((a != a) & (b != b))))


def test_cnv_oracle_reproduces_reference_state():
    g, prm = load()
    r = run_oracle(g, prm)
    for k in ("nblocks", "repeats", "chr_ave", "chr_sd", "rep_ave", "rep_sd", "rep_cnt", "biased", "blk_ave", "sample_blocks"):
        assert same(getattr(r, k), g["pre_" + k]), k
    assert int(g["pre_biased"]) == 3                                     # the thinned (AT)n runs take the biased-repeat path
    for k in ("win_sd", "win_cnt", "ave", "sd", "del_thr", "dup_thr", "windows", "n_high", "n_low"):
        assert same(getattr(r, k), g[k]), k
    assert np.array_equal(np.packbits(r.mask), g["mask_bits"])
    assert np.array_equal(r.z[::53], g["z_stride"]) and float(r.z.sum()) == float(g["z_sum"])
    assert float(np.abs(r.z).sum()) == float(g["z_abs_sum"]) and int((r.z != 0).sum()) == int(g["z_nonzero"])
    for mine, ref in ((r.dels, g["dels"]), (r.dups, g["dups"])):
        for f in ("start", "end", "z", "cn", "cs"):
            assert np.array_equal(mine[f], ref[f]), f
    assert r.vcf == str(g["vcf"]) and r.vcf.count("<DEL>") >= 1


def test_host_cnv_writer_reproduces_reference_records():
    g, prm = load()
    r = run_oracle(g, prm)
    calls = np.zeros(len(r.dels) + len(r.dups), dtype=CNV_CALL_DTYPE)
    for k, src in enumerate((r.dels, r.dups)):
        sl = slice(0, len(r.dels)) if k == 0 else slice(len(r.dels), None)
        calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
        calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
    assert hostlib.vcf_cnv(prm, "chra", calls) == str(g["vcf"])
    assert hostlib.vcf_cnv(prm, "chra", calls[:0]) == ""


def test_reference_bisections_match_python_except_pairs():
    """bisect_left/right of src/GROM.c:21630-21744 are true bisections except on two-element ranges (answer from the last element)."""
    import bisect, ctypes as C, itertools
    L = po.lib()
    for f in (L.oracle_bisect_left, L.oracle_bisect_right):
        f.restype = C.c_long; f.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long]
    for n in range(1, 8):
        for vals in itertools.combinations_with_replacement(range(4), n):
            a = np.array(vals, dtype=np.int32)
            for key in range(-1, 5):
                l = L.oracle_bisect_left(a.ctypes.data, key, 0, n); r = L.oracle_bisect_right(a.ctypes.data, key, 0, n)
                if n == 2:
                    assert l == (1 if key <= vals[1] else 2) and r == (1 if key < vals[1] else 2)
                else:
                    assert l == bisect.bisect_left(vals, key) and r == bisect.bisect_right(vals, key)
