"""End to end through the public API: BAM + FASTA in, VCF body out (grom_b200.pipeline) == the reference's own VCF for the committed
golden data set, with the library statistics measured from the BAM (find_insert_mean) instead of taken from the reference's cache."""
import gzip
import os

import numpy as np
import pytest

from util import GOLDEN
from grom_b200 import pipeline
from grom_b200.params import Params
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,rmdup", [("default", 0), ("rmdup", 1)])
def test_pipeline_reproduces_reference_vcf(tmp_path, tag, rmdup):
    g = np.load(os.path.join(GOLDEN, f"g1_{tag}.npz"))
    text, prm = pipeline.call_variants(os.path.join(GOLDEN, "g1.bam"), os.path.join(GOLDEN, "g1.fa.gz"), Params.default(rmdup=rmdup))
    m = g["mean"]
    assert (prm.insert_mean, prm.lseq, prm.insert_min, prm.insert_max) == (int(max(m[0], m[1])), int(m[1]), int(m[2]), int(m[3]))
    mine = "".join(text[t] for t in sorted(text)).splitlines(keepends=True)
    ref = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#")]
    assert len(ref) > 100 and po.normalise_records(mine) == po.normalise_records(ref)
    out = tmp_path / "out.vcf"
    pipeline.write_vcf(str(out), text)
    assert [l for l in open(out) if not l.startswith("#")] == mine
