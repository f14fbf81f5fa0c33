"""Binomial tables (reference src/GROM.c:21134-21586) pinned against the reference's own golden VCF and,
when the reference binary has written them, against its table files."""
import os
import re

import numpy as np
import pytest

from util import ROOT, GOLDEN, tables
from grom_b200 import hostlib

REF_VCF = "/root/reference/test_data/test_outuput_tilapia_SAMD00023995_GL831235-1.vcf"
LOCAL_SNV = os.path.join(GOLDEN, "tilapia_golden_snv.tsv.gz")


def _tilapia_snv_records():
    """(A,C,G,T, alt, PR text, AF text, GT) of every SNV record of the reference's golden tilapia VCF.
    A compact copy of just those columns is committed (the GPU box has no /root/reference)."""
    import gzip
    if os.path.exists(LOCAL_SNV):
        with gzip.open(LOCAL_SNV, "rt") as f:
            return [l.rstrip("\n").split("\t") for l in f]
    rows = []
    for l in open(REF_VCF):
        if l.startswith("#"):
            continue
        f = l.rstrip("\n").split("\t")
        if f[8].startswith("GT:PR:AF:A:C:G:T"):
            v = f[9].split(":")
            rows.append([v[3], v[4], v[5], v[6], f[4], v[1], v[2], v[0], f[3]])
    with gzip.open(LOCAL_SNV, "wt") as g:
        for r in rows:
            g.write("\t".join(r) + "\n")
    return rows


def test_mq_table_reproduces_every_golden_tilapia_snv_score():
    """PR = mq_table[A+C+G+T][alt count] printed with %e (src/GROM.c:11137-11146, 15082); AF = float ratio;
    GT = round(ratio*ploidy) ones (src/GROM.c:15057-15079).  18,099 records."""
    hez, mq = tables()
    rows = _tilapia_snv_records()
    assert len(rows) == 18099
    for a, c, g, t, alt, pr, af, gt, ref in rows:
        cnt = [int(a), int(c), int(g), int(t)]
        total = sum(cnt)
        k = cnt["ACGT".index(alt)]
        val = mq[1000][k * 1000 // total] if total > 1000 else mq[total][k]
        assert "%e" % val == pr, (cnt, alt, pr, val)
        ratio = float(np.float32(k) / np.float32(total))
        assert "%e" % ratio == af
        cn = int(np.floor(ratio * 2 + 0.5)) or 1
        assert gt == "/".join("1" if i < cn else "0" for i in range(2))
        assert ref.upper() != alt and k >= 3 and ratio >= 0.2


def test_tables_match_reference_written_files():
    d = os.path.join(ROOT, "oracle", "_ref")
    hp = os.path.join(d, "GROM_hez_binom_table_1000.txt")
    mp = os.path.join(d, "GROM_mq_binom_table_20_1000.txt")
    if not (os.path.exists(hp) and os.path.exists(mp)):
        pytest.skip("reference table files not generated in oracle/_ref")
    hez, mq = tables()
    lhez, lmq = hostlib.tables(d, 20)
    fmt = np.vectorize(lambda x: float("%e" % x))
    assert np.array_equal(fmt(hez), lhez)
    assert np.array_equal(fmt(mq), lmq)


def test_table_shape_properties():
    hez, mq = tables()
    assert mq[5][0] == 1.0 and np.all(np.diff(mq[200][:20]) <= 0)          # upper tail decreases (before the long factorial overflows)
    assert np.all(hez[:1000, 1000] == 1.0) and np.all(hez[0, :1000] == 1.0)  # row 0 / last column (src/GROM.c:21301-21316)
    assert hez[1000][0] == 1.0                                              # row 1000 is left untransformed
    pv, sd = hostlib.pval2sd()
    assert sd[0] == 10.0 and sd[-1] == 0.0 and np.all(np.diff(pv) >= 0) and abs(pv[-1] - 0.5) < 1e-6
