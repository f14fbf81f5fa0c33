"""BASELINE.json configs 1 and 2 (tilapia GL831235-1, default flags / -M).  The reference's own BAM for this contig is a missing blob, so
the fixture is a stand-in: 5x simulated pairs ON THE REFERENCE'S REAL FASTA (soft-masked lower case, 153 k N; tests/golden/
oreNil2_GL831235-1.fa.gz) and the VCF the reference's prebuilt binary wrote for them (tests/golden/make_tilapia_standin.py).
CPU: the reads regenerate bit-identically; oracle -> product host stages == the reference's records.  GPU: the C host program
(GROM_b200: BAM decode -> CUDA path -> record text) == the reference's records, default flags and -M, and the .ctx.vcf body."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import GOLDEN, ROOT, tables_7digit
from grom_b200 import hostlib
from grom_b200.params import CNV_CALL_DTYPE, Params
from oracle import pyoracle as po
from tools import synth

sys.path.insert(0, GOLDEN)
import make_tilapia_standin as mk  # noqa: E402

FA_GZ = os.path.join(GOLDEN, "oreNil2_GL831235-1.fa.gz")


@pytest.fixture(scope="module")
def standin():
    g = np.load(os.path.join(GOLDEN, "g6_tilapia_standin.npz"))
    chars = mk.load_fasta(FA_GZ)
    cs = mk.simulate(chars)
    return g, chars, cs


def test_fasta_is_the_reference_file_and_reads_regenerate(standin):
    g, chars, cs = standin
    assert len(chars) == 2_653_313
    assert int((chars == ord("N")).sum()) > 150_000 and int(np.isin(chars, np.frombuffer(b"acgt", dtype=np.uint8)).sum()) > 100_000
    assert mk.reads_digest(cs) == str(g["digest"]) and [c.batch.n_reads for c in cs] == list(g["n_reads"])


def _calls(cn):
    calls = np.zeros(len(cn.dels) + len(cn.dups), dtype=CNV_CALL_DTYPE)
    for k, src in enumerate((cn.dels, cn.dups)):
        sl = slice(0, len(cn.dels)) if k == 0 else slice(len(cn.dels), None)
        calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
        calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
    return calls


@pytest.mark.parametrize("tag,rmdup", [("default", 0), ("rmdup", 1)])
def test_oracle_and_host_stages_reproduce_the_reference_records(standin, tag, rmdup):
    g, chars, cs = standin
    m = g["mean"]
    prm = Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]), rmdup=rmdup)
    hez, mq = tables_7digit()
    ref = str(g[f"vcf_{tag}"]).splitlines(keepends=True)
    mine, recs = [], []
    for tid, c in enumerate(cs):
        name = c.name.lower()
        r = po.run_chr(prm, c.batch, c.chars, hez, mq)
        cn = po.cnv_run(prm, name, c.chars, r["gc"], r["acgt"], r["rd_mq"], r["rd_rd"], r["rd_low"])
        mine += hostlib.vcf_contig(prm, name, c.chars, r.snv, r.snv_ave_rd, r.ins, r.del_ev, r.sv_ev, _calls(cn)).splitlines(keepends=True)
        recs.append(hostlib.ctx_contig(prm, tid, r.sv_ev))
    assert len(ref) > 1000 and po.normalise_records(mine) == po.normalise_records(ref)
    assert hostlib.ctx_vcf(prm, [c.name for c in cs], np.concatenate(recs)) == str(g[f"ctx_{tag}"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,flags", [("default", []), ("rmdup", ["-M"])])
def test_c_host_program_reproduces_the_reference_records(standin, tmp_path, tag, flags):
    g, chars, cs = standin
    fa, bam = synth.write_dataset(str(tmp_path / "til"), cs)
    exe = os.path.join(ROOT, "grom_b200", "GROM_b200")
    r = subprocess.run([exe, "-i", bam, "-r", fa, "-o", str(tmp_path / "o.vcf")] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    m = g["mean"]
    assert f"insert mean, insert minimum, insert maximum: {int(max(m[0], m[1]))} {int(m[2])} {int(m[3])}" in r.stdout
    mine = [l for l in open(tmp_path / "o.vcf") if not l.startswith("#")]
    ref = str(g[f"vcf_{tag}"]).splitlines(keepends=True)
    assert len(ref) > 1000 and po.normalise_records(mine) == po.normalise_records(ref)
    assert "".join(l for l in open(tmp_path / "o.ctx.vcf") if not l.startswith("#")) == str(g[f"ctx_{tag}"])
