"""Host stage of the scan (grom_b200/host/vcf.c): emission filters, small-deletion pairing state machine and VCF text,
checked against the reference's own records (golden fixture) using the oracle's candidates as input."""
import os

import numpy as np
import pytest

from util import GOLDEN, golden_batches, golden_params, load_golden_fasta, tables_7digit
from grom_b200 import hostlib
from oracle import pyoracle as po


@pytest.mark.parametrize("tag,rmdup", [("default", 0), ("rmdup", 1)])
def test_host_writer_reproduces_reference_records(tag, rmdup):
    names, batches = golden_batches()
    fasta = load_golden_fasta()
    hez, mq = tables_7digit()
    g = np.load(os.path.join(GOLDEN, f"g1_{tag}.npz"))
    prm = golden_params(g, rmdup)
    vcf = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#")]
    n_ref = n_mine = 0
    for tid, name in enumerate(names):
        n = name.lower()
        r = po.run_chr(prm, batches[tid], fasta[name], hez, mq)
        snv = hostlib.vcf_snv(prm, n, fasta[name], r.snv, r.snv_ave_rd).splitlines(keepends=True)
        ins = hostlib.vcf_ins(prm, n, fasta[name], r.ins).splitlines(keepends=True)
        dele = hostlib.vcf_smalldel(prm, n, fasta[name], r.del_ev).splitlines(keepends=True)
        mine = [l for l in vcf if l.startswith(n + "\t")]
        assert snv == [l for l in mine if l.split("\t")[2] == ""]
        assert po.normalise_records(ins) == po.normalise_records([l for l in mine if "\tSPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP\t" in l])
        assert dele == [l for l in mine if "\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT:SSC:ESC:HP\t" in l]
        # every record of the contig through the one-call writer, with the read-depth calls of the CNV oracle
        cn = po.cnv_run(prm, n, fasta[name], r["gc"], r["acgt"], r["rd_mq"], r["rd_rd"], r["rd_low"])
        from grom_b200.params import CNV_CALL_DTYPE
        calls = np.zeros(len(cn.dels) + len(cn.dups), dtype=CNV_CALL_DTYPE)
        for k, src in enumerate((cn.dels, cn.dups)):
            sl = slice(0, len(cn.dels)) if k == 0 else slice(len(cn.dels), None)
            calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
            calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
        full = hostlib.vcf_contig(prm, n, fasta[name], r.snv, r.snv_ave_rd, r.ins, r.del_ev, r.sv_ev, calls)
        assert po.normalise_records(full.splitlines(keepends=True)) == po.normalise_records(mine)
        # the reference prints the classes in this order per contig: SNV ... small INS, small DEL ...
        n_ref += len(mine); n_mine += len(snv) + len(ins) + len(dele)
    assert n_mine >= 0.95 * n_ref          # the remaining records are the structural classes, covered by the full-contig comparison above


def test_smalldel_state_machine_edge_cases():
    from grom_b200.params import DEL_EVENT_DTYPE, Params
    prm = Params.default()
    fa = np.frombuffer(b"ACGT" * 2000, dtype=np.uint8).copy()
    ev = np.zeros(0, dtype=DEL_EVENT_DTYPE)
    assert hostlib.vcf_smalldel(prm, "c", fa, ev) == ""
    # one complete pair followed by a second start: only the first is emitted (the loop stops before the last entry)
    ev = np.zeros(3, dtype=DEL_EVENT_DTYPE)
    ev["pos"] = [1000, 1004, 3000]; ev["kind"] = [0, 1, 0]; ev["pr"] = 1e-9; ev["weight"] = 60; ev["rd"] = 20; ev["rdist"] = 5
    out = hostlib.vcf_smalldel(prm, "c", fa, ev).splitlines()
    assert len(out) == 1 and out[0].startswith("c\t1001\t.\t" + bytes(fa[1000:1005]).decode() + "\t.\t.\t.\tEND=1005\t")


def test_contig_writer_all_record_classes_against_reference_vcf():
    """tests/golden/g4_vcf.npz: the inputs of gromhost_vcf_contig for a contig on which the reference (prebuilt dist/GROM) emits every
    record class, and the reference's own text: 1,695 records incl. <DUP>, <INV>, <INS>, paired-end and read-depth <DEL>."""
    g = np.load(os.path.join(GOLDEN, "g4_vcf.npz"))
    m = g["mean"]
    from grom_b200.params import Params
    prm = Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]))
    mine = hostlib.vcf_contig(prm, "chra", g["fasta"], g["snv"], float(g["snv_ave_rd"]), g["ins"], g["del_ev"], g["sv_ev"], g["cnv"])
    ref = str(g["vcf"]).splitlines(keepends=True)
    assert po.normalise_records(mine.splitlines(keepends=True)) == po.normalise_records(ref)
    for tag in ("<DUP>", "<INV>", "<INS>", "<DEL>", "SD:Z:CN:CS", "SSC:ESC:HP", "SSC:HP"):
        assert any(tag in l for l in ref), tag
    # per-class writers agree with the one-call writer on their share
    assert hostlib.vcf_snv(prm, "chra", g["fasta"], g["snv"], float(g["snv_ave_rd"])) == "".join(l for l in ref if l.split("\t")[2] == "")
    assert hostlib.vcf_cnv(prm, "chra", g["cnv"]) == "".join(l for l in ref if "\tSD:Z:CN:CS\t" in l)


def test_snv_records_formatted_by_all_threads_equal_the_serial_text(monkeypatch):
    """From a few thousand candidates on, gromhost_vcf_snv formats contiguous shares of them on all threads and lays the pieces end to end
    (printf's %e costs microseconds per record and a chromosome has tens of thousands): same bytes as one thread, also when the shares are
    uneven, when most candidates are filtered out, and through the one-call contig writer.  Forced here with GROMHOST_SNV_PAR_MIN."""
    from grom_b200.params import Params
    g = np.load(os.path.join(GOLDEN, "g4_vcf.npz"))
    m = g["mean"]
    prm = Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]))
    fa, ave = g["fasta"], float(g["snv_ave_rd"])
    base = g["snv"]
    assert len(base) > 200
    snv = np.concatenate([base] * 9)[: 8 * len(base) + 37].copy()       # ~10 k candidates, not a multiple of anything
    snv["pr"][::7] *= 1e-200                                             # the slow corner of %e
    snv["ratio"][5::11] = 0.0                                            # ... and candidates the depth filter drops, in runs of different lengths per share
    snv["v"][5::11] = snv["v"][5::11] + 1000
    for par_min, threads in (("1", None), ("1", "3"), ("100000000", None)):
        monkeypatch.setenv("GROMHOST_SNV_PAR_MIN", par_min)
        if threads:
            monkeypatch.setenv("OMP_NUM_THREADS", threads)
        text = hostlib.vcf_snv(prm, "chra", fa, snv, ave)
        if par_min == "1" and threads is None:
            first = text
        assert text == first
    assert first.count("\n") > 0.8 * len(snv) * 0.9 and "e-2" in first
    monkeypatch.setenv("GROMHOST_SNV_PAR_MIN", "1")
    a = hostlib.vcf_contig(prm, "chra", fa, g["snv"], ave, g["ins"], g["del_ev"], g["sv_ev"], g["cnv"])
    monkeypatch.setenv("GROMHOST_SNV_PAR_MIN", "100000000")
    assert a == hostlib.vcf_contig(prm, "chra", fa, g["snv"], ave, g["ins"], g["del_ev"], g["sv_ev"], g["cnv"])


TILAPIA_OUT = "/root/reference/test_data/test_outuput_tilapia_SAMD00023995_GL831235-1"


@pytest.mark.skipif(not os.path.exists(TILAPIA_OUT + ".vcf"), reason="the reference's example output (build container only)")
def test_header_blocks_equal_the_reference_example_output(tmp_path):
    """gromhost_vcf_header (used by the C host program and by grom_b200.pipeline.write_vcf / write_ctx_vcf) against the header lines of the
    two files the reference ships as its example output, date and reference path aside."""
    from grom_b200 import pipeline
    for ext, is_ctx in ((".vcf", False), (".ctx.vcf", True)):
        ref = [l for l in open(TILAPIA_OUT + ext) if l.startswith("#")]
        mine = hostlib.vcf_header("some/ref.fa", is_ctx).splitlines(keepends=True)
        keep = lambda ls: [l for l in ls if not l.startswith(("##fileDate=", "##reference="))]
        assert len(ref) > 20 and keep(mine) == keep(ref)
        assert mine[1].startswith("##fileDate=") and mine[2] == "##reference=some/ref.fa\n"
    pipeline.write_vcf(str(tmp_path / "o.vcf"), {1: "b\n", 0: "a\n"}, "r.fa")
    pipeline.write_ctx_vcf(str(tmp_path / "o.ctx.vcf"), "c\n", "r.fa")
    assert open(tmp_path / "o.vcf").read() == hostlib.vcf_header("r.fa", False) + "a\nb\n"
    assert open(tmp_path / "o.ctx.vcf").read() == hostlib.vcf_header("r.fa", True) + "c\n"
