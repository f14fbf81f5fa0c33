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
    # one contig at a time == three contigs in flight (threads / streams; the golden BAM has three contigs)
    serial, _ = pipeline.call_variants(os.path.join(GOLDEN, "g1.bam"), os.path.join(GOLDEN, "g1.fa.gz"), Params.default(rmdup=rmdup), lanes=1)
    assert serial == text and len(text) == 3


def test_pipeline_translocations(tmp_path):
    """BAM + FASTA with reciprocal inter-contig clusters: the .ctx.vcf body from the GPU pipeline == the one the host stages produce from
    the oracle's gate events (which tests/test_sv_lists_golden.py pins on the reference's own .ctx.vcf)."""
    from util import tables_7digit
    from grom_b200 import hostlib
    from tools import synth
    spec = synth.SynthSpec(contigs=[("chrA", 300_000), ("chrB", 300_000), ("chrZ", 50_000)], depth=30, seed=31, sv_classes=20, disc_frac=0.005)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "ctx"), cs)
    ctx = {}
    text, prm = pipeline.call_variants(bam, fa, Params.default(), ctx_out=ctx)
    body = pipeline.ctx_vcf_text(prm, [c.name for c in cs], ctx)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    want = []
    with hostlib.Bam(bam) as b:
        for tid, c in enumerate(cs):
            r = po.run_chr(prm, b.read_target(tid), c.chars, hez, mq)
            want.append(hostlib.ctx_contig(prm, tid, r.sv_ev))
    assert body == hostlib.ctx_vcf(prm, [c.name for c in cs], np.concatenate(want)) and body.count("SVTYPE=BND") >= 4
