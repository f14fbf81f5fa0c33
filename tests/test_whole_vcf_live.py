"""Live whole-body check on fresh data (build container only): BAM + FASTA -> the oracle's evidence, gates and read-depth calls -> the
PRODUCT host stages (list builder, pairing, filters, record writer: grom_b200/host) == the reference's VCF body, every contig and
every record class, under flag combinations no fixture holds (-M, -q/-b, -v, read length).  The GPU twin of this test
(tests/test_gpu_pipeline.py) swaps the oracle for the CUDA path on the committed data set.
Skipped where oracle/_ref/GROM_ref is absent."""
import numpy as np
import pytest

from grom_b200 import hostlib
from grom_b200.params import CNV_CALL_DTYPE, Params
from oracle import pyoracle as po
from tools import synth

pytestmark = pytest.mark.skipif(not po.have_reference("ref"), reason="oracle/_ref/GROM_ref not built")


def _calls(cn):
    calls = np.zeros(len(cn.dels) + len(cn.dups), dtype=CNV_CALL_DTYPE)
    for k, src in enumerate((cn.dels, cn.dups)):
        sl = slice(0, len(cn.dels)) if k == 0 else slice(len(cn.dels), None)
        calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
        calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
    return calls


@pytest.mark.parametrize("seed,read_len,flags", [
    (41, 150, dict()),
    (42, 100, dict(rmdup=1, min_mapq=30, min_base_qual=26)),
    (43, 150, dict(rmdup=1, pval_threshold=0.01)),
])
def test_whole_vcf_body_against_live_reference(tmp_path, seed, read_len, flags):
    spec = synth.SynthSpec(contigs=[("chrA", 700_000), ("chrB", 150_000), ("chrZ", 50_000)], depth=40, seed=seed, read_len=read_len,
                           ins_mean=400.0 * read_len / 150, ins_sd=40, ins_floor=read_len + 20, dup_frac=0.05, sa_frac=0.5, disc_frac=0.03,
                           sv_sites_per_mb=12.0, munmap_frac=0.01, sv_classes=12, cnv_per_mb=1.5, cnv_min=15000, cnv_max=30000)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), cs)
    args = (["-M"] if flags.get("rmdup") else [])
    if "min_mapq" in flags:
        args += ["-q", flags["min_mapq"], "-b", flags["min_base_qual"]]
    if "pval_threshold" in flags:
        args += ["-v", flags["pval_threshold"]]
    po.run_reference(bam, fa, str(tmp_path / "o.vcf"), args=args, seed=1)
    m = po.read_mean_file(bam)
    kw = dict(flags)
    if "min_mapq" in kw:
        kw["rd_min_mapq"] = kw["min_mapq"]                    # src/GROM.c:22102
    if "pval_threshold" in kw:
        kw["pval_threshold1"] = kw["pval_threshold"]          # src/GROM.c:22101
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"],
                         lseq=m["lseq"], **kw)
    hez, mq = po.reference_tables(prm.min_mapq)
    body = [l for l in open(str(tmp_path / "o.vcf")) if not l.startswith("#")]
    kinds = set()
    with hostlib.Bam(bam) as b:
        for tid, c in enumerate(cs):
            name = c.name.lower()
            r = po.run_chr(prm, b.read_target(tid), c.chars, hez, mq)
            cn = po.cnv_run(prm, name, c.chars, r["gc"], r["acgt"], r["rd_mq"], r["rd_rd"], r["rd_low"])
            mine = hostlib.vcf_contig(prm, name, c.chars, r.snv, r.snv_ave_rd, r.ins, r.del_ev, r.sv_ev, _calls(cn)).splitlines(keepends=True)
            ref = [l for l in body if l.startswith(name + "\t")]
            assert po.normalise_records(mine) == po.normalise_records(ref), (name, len(mine), len(ref))
            for l in ref:
                f = l.split("\t")
                kinds.add(f[4] + ("/cnv" if f[8] == "SD:Z:CN:CS" else "") if f[4].startswith("<") else ("snv" if f[2] == "" else "indel"))
    assert {"snv", "indel", "<DEL>"} <= kinds and len(kinds) >= 5, kinds


@pytest.mark.parametrize("seed,rmdup,n_contigs", [(51, 0, 3), (52, 1, 4)])
def test_translocation_records_against_live_reference(tmp_path, seed, rmdup, n_contigs):
    """<out>.ctx.vcf: per-contig candidate merge + filter (gromhost_ctx_contig) and genome-level mate pairing (gromhost_ctx_vcf,
    src/GROM.c:22400-22770) on fresh data with reciprocal inter-contig clusters; three and four contigs, with and without -M."""
    contigs = [("chrA", 300_000), ("chrB", 300_000), ("chrC", 200_000), ("chrZ", 50_000)]
    contigs = contigs[:n_contigs - 1] + contigs[-1:]
    spec = synth.SynthSpec(contigs=contigs, depth=30, seed=seed, sv_classes=20, disc_frac=0.005, dup_frac=0.05 if rmdup else 0.0)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), cs)
    po.run_reference(bam, fa, str(tmp_path / "o.vcf"), args=(["-M"] if rmdup else []))
    ref = [l for l in open(str(tmp_path / "o.ctx.vcf")) if not l.startswith("#")]
    m = po.read_mean_file(bam)
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"],
                         lseq=m["lseq"], rmdup=rmdup)
    hez, mq = po.reference_tables(20)
    recs = []
    with hostlib.Bam(bam) as b:
        for tid, c in enumerate(cs):
            r = po.run_chr(prm, b.read_target(tid), c.chars, hez, mq)
            recs.append(hostlib.ctx_contig(prm, tid, r.sv_ev))
    assert hostlib.ctx_vcf(prm, [c.name for c in cs], np.concatenate(recs)).splitlines(keepends=True) == ref
    assert len(ref) >= 4 and all("SVTYPE=BND" in l for l in ref)
