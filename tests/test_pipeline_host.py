"""The Python driver's host side without a GPU: grom_b200.pipeline.call_variants with the CUDA binding replaced -- in this test only -- by
a stand-in that answers from the CPU oracle.  What is checked is the driver itself: the statistics pass straight over the BAM, contigs
loaded one at a time through the C FASTA reader, the batcher's own memory handed over (OwnedBatch), lanes / rebind / admission control,
per-contig text and the translocation records; the result must be the reference's VCF body for the committed data set.  (The GPU twin,
tests/test_gpu_pipeline.py, runs the same driver on the CUDA library.)"""
import gzip
import os
import types

import numpy as np
import pytest

from util import GOLDEN
from grom_b200 import hostlib, pipeline
from grom_b200.params import CNV_CALL_DTYPE, Params
from oracle import pyoracle as po
from tools import synth


def _calls(cn):
    calls = np.zeros(len(cn.dels) + len(cn.dups), dtype=CNV_CALL_DTYPE)
    for k, src in enumerate((cn.dels, cn.dups)):
        sl = slice(0, len(cn.dels)) if k == 0 else slice(len(cn.dels), None)
        calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
        calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
    return calls


class OracleGpu:
    """Stand-in for grom_b200.gpu with the calls the driver makes; every answer comes from oracle/ (test infrastructure)."""

    def __init__(self):
        self.log = []
        outer = self

        class Chromosome:
            def __init__(self, tid, chars, stream=None):
                self.tid, self.chars, self.cap, self.batch = tid, np.array(chars), len(chars), None
                outer.log.append(("begin", tid))

            def rebind(self, tid, chars):
                if len(chars) > self.cap:
                    return False
                self.tid, self.chars, self.batch = tid, np.array(chars), None
                outer.log.append(("rebind", tid))
                return True

            def push_reads(self, batch):
                assert isinstance(batch, hostlib.OwnedBatch) and batch.as_c().n_reads == batch.n_reads
                piece = batch.to_numpy()                           # the stand-in copies; the CUDA library uploads
                outer.log.append(("push", self.tid, piece.n_reads))
                if self.batch is not None:                         # consecutive pieces of one target (gromgpu_push_reads may be called repeatedly)
                    assert piece.pos[0] >= self.batch.pos[-1]
                    piece = synth.concat_batches([self.batch, piece])
                self.batch = piece

            def sync(self):
                pass

            def finish(self):
                self.r = po.run_chr(outer.prm, self.batch, self.chars, outer.hez, outer.mq)
                return self.r

            def cnv(self, params=None):
                r = self.r
                cn = po.cnv_run(params, outer.names[self.tid].lower(), self.chars, r["gc"], r["acgt"], r["rd_mq"], r["rd_rd"], r["rd_low"])
                return types.SimpleNamespace(calls=_calls(cn))

            def close(self):
                outer.log.append(("close", self.tid))
        self.Chromosome = Chromosome

    def init(self, device, hez, mq, prm):
        self.hez, self.mq, self.prm = hez, mq, prm

    def stream_create(self):
        return 1

    def stream_destroy(self, s):
        pass

    def device_free_bytes(self):
        return 1 << 40

    def chr_bytes_estimate(self, n_chars, n_reads, n_slots):
        assert n_reads > 0 and n_slots >= n_reads
        return 1 << 20


@pytest.mark.parametrize("tag,rmdup,lanes", [("default", 0, 3), ("rmdup", 1, 1)])
def test_driver_reproduces_the_reference_vcf_with_the_oracle_behind_it(tmp_path, monkeypatch, tag, rmdup, lanes):
    g = np.load(os.path.join(GOLDEN, f"g1_{tag}.npz"))
    fa = tmp_path / "g1.fa"
    fa.write_bytes(gzip.open(os.path.join(GOLDEN, "g1.fa.gz"), "rb").read())          # plain file: the C FASTA reader
    fake = OracleGpu()
    with hostlib.Bam(os.path.join(GOLDEN, "g1.bam")) as b:
        fake.names = list(b.names)
    monkeypatch.setattr(pipeline, "gpu", fake)
    ctx = {}
    text, prm = pipeline.call_variants(os.path.join(GOLDEN, "g1.bam"), str(fa), Params.default(rmdup=rmdup), ctx_out=ctx, lanes=lanes)
    m = g["mean"]
    assert (prm.insert_mean, prm.lseq, prm.insert_min, prm.insert_max) == (int(max(m[0], m[1])), int(m[1]), int(m[2]), int(m[3]))
    mine = "".join(text[t] for t in sorted(text)).splitlines(keepends=True)
    ref = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#")]
    assert len(ref) > 100 and po.normalise_records(mine) == po.normalise_records(ref)
    assert sorted(ctx) == sorted(text) and len(text) == 3
    # handles: one per lane, begun for the largest contig a lane sees and rebound for the rest
    begun = [e for e in fake.log if e[0] == "begin"]
    assert len(begun) == lanes and len([e for e in fake.log if e[0] == "close"]) == lanes
    if lanes == 1:
        assert [e[0] for e in fake.log if e[0] != "push"] == ["begin", "rebind", "rebind", "close"]
    # the same with every contig decoded and pushed in pieces of >= 500 reads (small decode windows so that the pieces are small)
    monkeypatch.setenv("GROMHOST_WINDOW_BLOCKS", "2")
    fake.log.clear()
    sliced, _ = pipeline.call_variants(os.path.join(GOLDEN, "g1.bam"), str(fa), Params.default(rmdup=rmdup), lanes=lanes, slice_reads=500)
    assert sliced == text
    pushes = [e for e in fake.log if e[0] == "push"]
    assert len(pushes) > 6 and sum(e[2] for e in pushes) == 8000 + 5000 + 1600


def test_driver_reports_a_fasta_that_does_not_match_the_bam(tmp_path, monkeypatch):
    data = gzip.open(os.path.join(GOLDEN, "g1.fa.gz"), "rb").read()
    lines = data.split(b"\n")
    del lines[1]                                                                       # first contig one line short
    fa = tmp_path / "short.fa"
    fa.write_bytes(b"\n".join(lines))
    fake = OracleGpu()
    with hostlib.Bam(os.path.join(GOLDEN, "g1.bam")) as b:
        fake.names = list(b.names)
    monkeypatch.setattr(pipeline, "gpu", fake)
    with pytest.raises(ValueError, match="in the BAM header"):
        pipeline.call_variants(os.path.join(GOLDEN, "g1.bam"), str(fa), Params.default(), lanes=2)
