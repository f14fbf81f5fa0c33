"""Regenerates the committed golden fixtures by running the REFERENCE ITSELF (oracle/_ref/GROM_ref, the
reference's own translation unit built with dump hooks, and oracle/_ref/GROM_dist, its prebuilt binary)
on a small synthetic data set.  Run in the build container (needs /root/reference for `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/g1.fa.gz, g1.bam, g1.bam.bai      the input (so nothing depends on numpy's RNG stream)
    tests/golden/g1_default.npz, g1_rmdup.npz      reference arrays at every scanned position, per-read keep
                                                   flags, CNV depth arrays, VCF records, library statistics
    tests/golden/g2_cnv.npz                        read-depth CNV path of a 1.2 Mb contig with a planted one-copy loss and thinned
                                                   (AT)n runs: the reference's own CNV depth / GC arrays as compact input, and its
                                                   pre-statistics, per-bin distributions, mask, z list (sampled + checksums), window
                                                   sd table, calls with copy number, and VCF records as expected output
"""
import gzip
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

CONTIGS = [("chrG", 40_000), ("chrH", 25_000), ("chrZ", 8_000)]


def main():
    spec = synth.SynthSpec(contigs=CONTIGS, depth=30, seed=2024, dup_frac=0.06, clip_frac=0.04, disc_frac=0.02,
                           snv_every=400, indel_every=3000, hardclip_frac=0.005, refskip_frac=0.002, long_name_frac=0.01)
    cs = synth.simulate(spec)
    tmp = tempfile.mkdtemp()
    fa, bam = synth.write_dataset(os.path.join(tmp, "g1"), cs)
    shutil.copy(bam, os.path.join(HERE, "g1.bam"))
    shutil.copy(bam + ".bai", os.path.join(HERE, "g1.bam.bai"))
    with open(fa, "rb") as f, gzip.open(os.path.join(HERE, "g1.fa.gz"), "wb", 9) as g:
        g.write(f.read())
    for tag, args in (("default", []), ("rmdup", ["-M"])):
        dump = os.path.join(tmp, "dump_" + tag)
        vcf = os.path.join(tmp, tag + ".vcf")
        po.run_reference(bam, fa, vcf, args=args, dump_dir=dump, kind="ref")
        vcf_dist = os.path.join(tmp, tag + ".dist.vcf")
        po.run_reference(bam, fa, vcf_dist, args=args, kind="dist")
        rec = [l for l in open(vcf) if not l.startswith("##")]
        rec_dist = [l for l in open(vcf_dist) if not l.startswith("##")]
        assert po.normalise_records(rec) == po.normalise_records(rec_dist), "white-box build and prebuilt reference binary disagree"
        rec = rec_dist      # keep the prebuilt binary's text (identical up to the uninitialised ECO/EOT fields)
        out = dict(vcf=np.array("".join(rec)), mean=np.array([po.read_mean_file(bam)[k] for k in
                                                              ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")]))
        for name, _ in CONTIGS:
            n = name.lower()
            sd = po.load_scan_dump(dump, n)
            out[f"{n}_scan_pos"] = sd["pos"]
            out[f"{n}_scan_v"] = sd["v"]
            out[f"{n}_scan_d"] = sd["d"]
            out[f"{n}_reads"] = po.load_reads_dump(dump, n)
            out[f"{n}_depth"] = po.load_depth_dump(dump, n)
            out[f"{n}_gc"] = po.load_gc_dump(dump, n)
        np.savez_compressed(os.path.join(HERE, f"g1_{tag}.npz"), **out)
        print(tag, "records", len(rec))
    shutil.rmtree(tmp)


def main_cnv():
    spec = synth.SynthSpec(contigs=[("chrA", 1_200_000), ("chrZ", 50_000)], depth=30, seed=21, cnv_per_mb=0.9, disc_frac=0.005,
                           sv_sites_per_mb=1.0, low_mapq_frac=0.05, at_repeats=260, cnv_min=15000, cnv_max=30000)
    cs = synth.simulate(spec)
    tmp = tempfile.mkdtemp()
    fa, bam = synth.write_dataset(os.path.join(tmp, "g2"), cs)
    dump = os.path.join(tmp, "dump")
    vcf = os.path.join(tmp, "g2.vcf")
    po.run_reference(bam, fa, vcf, dump_dir=dump, kind="ref", seed=5)
    vcf_dist = os.path.join(tmp, "g2.dist.vcf")
    po.run_reference(bam, fa, vcf_dist, kind="dist")
    cnv_lines = [l for l in open(vcf) if l.startswith("chra\t") and "\tSD:Z:CN:CS\t" in l]
    assert cnv_lines == [l for l in open(vcf_dist) if l.startswith("chra\t") and "\tSD:Z:CN:CS\t" in l], "white-box and prebuilt binary disagree"
    assert len(cnv_lines) >= 1
    m = po.read_mean_file(bam)
    dep = po.load_depth_dump(dump, "chra")
    gcd = po.load_gc_dump(dump, "chra")
    pre = po.load_cnvpre_dump(dump, "chra")
    d = po.load_cnv_dump(dump, "chra")
    assert dep[1].max() < 256 and dep[2].max() < 256 and dep[0].max() < 65536 and pre["biased"] != -1
    out = dict(fasta=cs[0].chars, rd_mq=dep[0].astype(np.uint16), rd_rd=dep[1].astype(np.uint8), rd_low=dep[2].astype(np.uint8),
               gc=gcd[0].astype(np.uint8), acgt=gcd[1].astype(np.uint8),
               mean=np.array([m[k] for k in ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")]), seed=np.array(5),
               vcf=np.array("".join(cnv_lines)), z_stride=d["z"][::53].copy(), z_sum=np.array(float(d["z"].sum())),
               z_abs_sum=np.array(float(np.abs(d["z"]).sum())), z_nonzero=np.array(int((d["z"] != 0).sum())), mask_bits=np.packbits(d["mask"]))
    for k in ("nblocks", "repeats", "chr_ave", "chr_sd", "rep_ave", "rep_sd", "rep_cnt", "biased", "blk_ave", "sample_blocks"):
        out["pre_" + k] = np.asarray(pre[k])
    for k in ("win_sd", "win_cnt", "ave", "sd", "del_thr", "dup_thr", "windows", "n_high", "n_low", "dels", "dups"):
        out[k] = d[k]
    np.savez_compressed(os.path.join(HERE, "g2_cnv.npz"), **out)
    print("g2_cnv: records", len(cnv_lines), "dels", len(d["dels"]), "dups", len(d["dups"]), "biased", pre["biased"],
          "size %.2f MB" % (os.path.getsize(os.path.join(HERE, "g2_cnv.npz")) / 1e6))
    shutil.rmtree(tmp)


def main_sv():
    """g3_svlists.npz: candidate lists of the reference at the end of the per-position scan (hook at src/GROM.c:15164) for (a) the
    committed g1 BAM and (b) a data set with planted clusters of every structural-variant class, stored together with the gate
    events (the input of the host list builder) so the fixture does not depend on numpy's RNG stream."""
    from grom_b200 import hostlib
    from grom_b200.params import Params
    tmp = tempfile.mkdtemp()
    out = {}
    # (a) g1
    fa = os.path.join(tmp, "g1.fa")
    with gzip.open(os.path.join(HERE, "g1.fa.gz"), "rb") as f, open(fa, "wb") as g:
        g.write(f.read())
    bam = os.path.join(tmp, "g1.bam")
    shutil.copy(os.path.join(HERE, "g1.bam"), bam); shutil.copy(os.path.join(HERE, "g1.bam.bai"), bam + ".bai")
    dump = os.path.join(tmp, "dump_g1")
    po.run_reference(bam, fa, os.path.join(tmp, "g1.vcf"), dump_dir=dump, kind="ref")
    for name, _ in CONTIGS:
        for k, v in po.load_svlist_dump(dump, name.lower()).items():
            out[f"g1_{name.lower()}_{k}"] = v
    # (b) all classes
    spec = synth.SynthSpec(contigs=[("chrA", 400_000), ("chrB", 150_000), ("chrZ", 50_000)], depth=30, seed=14, dup_frac=0.05, sa_frac=0.5,
                           disc_frac=0.03, sv_sites_per_mb=10.0, munmap_frac=0.01, sv_classes=25)
    cs = synth.simulate(spec)
    fa2, bam2 = synth.write_dataset(os.path.join(tmp, "g3"), cs)
    dump2 = os.path.join(tmp, "dump_g3")
    po.run_reference(bam2, fa2, os.path.join(tmp, "g3.vcf"), dump_dir=dump2, kind="ref")
    m = po.read_mean_file(bam2)
    out["g3_mean"] = np.array([m[k] for k in ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")])
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"], lseq=m["lseq"])
    hez, mq = po.reference_tables(20)
    with hostlib.Bam(bam2) as b:
        for tid, c in enumerate(cs[:2]):
            r = po.run_chr(prm, b.read_target(tid), c.chars, hez, mq)
            out[f"g3_{c.name.lower()}_events"] = r.sv_ev
            ref = po.load_svlist_dump(dump2, c.name.lower())
            mine = po.normalise_sv_lists(hostlib.sv_lists(prm, r.sv_ev))
            for k, v in ref.items():
                assert v.tobytes() == mine[k].tobytes(), (c.name, k)
                out[f"g3_{c.name.lower()}_{k}"] = v
            print(c.name, {k: len(v) for k, v in ref.items()})
    np.savez_compressed(os.path.join(HERE, "g3_svlists.npz"), **out)
    print("g3_svlists: size %.2f MB" % (os.path.getsize(os.path.join(HERE, "g3_svlists.npz")) / 1e6))
    shutil.rmtree(tmp)


def main_vcf():
    """g4_vcf.npz: every input of the per-contig record writer (candidates, events, read-depth calls, FASTA) for a contig on which the
    reference emits every record class (SNV, <DUP>, <INV>, <INS>, small INS / DEL, <DEL>, read-depth <DEL>/<DUP>), and its VCF text."""
    from grom_b200 import hostlib
    from grom_b200.params import CNV_CALL_DTYPE, Params
    spec = synth.SynthSpec(contigs=[("chrA", 1_500_000), ("chrB", 200_000), ("chrZ", 50_000)], depth=40, seed=17, dup_frac=0.05, sa_frac=0.5,
                           disc_frac=0.03, sv_sites_per_mb=10.0, munmap_frac=0.01, sv_classes=12, cnv_per_mb=0.7, cnv_min=15000, cnv_max=30000)
    cs = synth.simulate(spec)
    tmp = tempfile.mkdtemp()
    fa, bam = synth.write_dataset(os.path.join(tmp, "g4"), cs)
    vcf = os.path.join(tmp, "g4.vcf")
    po.run_reference(bam, fa, vcf, kind="dist", seed=1)
    m = po.read_mean_file(bam)
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"], lseq=m["lseq"])
    hez, mq = po.reference_tables(20)
    c = cs[0]
    with hostlib.Bam(bam) as b:
        r = po.run_chr(prm, b.read_target(0), c.chars, hez, mq)
    cn = po.cnv_run(prm, "chra", c.chars, r["gc"], r["acgt"], r["rd_mq"], r["rd_rd"], r["rd_low"])
    calls = np.zeros(len(cn.dels) + len(cn.dups), dtype=CNV_CALL_DTYPE)
    for k, src in enumerate((cn.dels, cn.dups)):
        sl = slice(0, len(cn.dels)) if k == 0 else slice(len(cn.dels), None)
        calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
        calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
    ref = [l for l in open(vcf) if l.startswith("chra\t")]
    mine = hostlib.vcf_contig(prm, "chra", c.chars, r.snv, r.snv_ave_rd, r.ins, r.del_ev, r.sv_ev, calls).splitlines(keepends=True)
    assert po.normalise_records(mine) == po.normalise_records(ref)
    kinds = {}
    for l in ref:
        f = l.split("\t")
        k = f[4] + ("/cnv" if f[8] == "SD:Z:CN:CS" else "") if f[4].startswith("<") else ("snv" if f[2] == "" else "indel")
        kinds[k] = kinds.get(k, 0) + 1
    print("g4 record classes:", kinds)
    assert all(k in kinds for k in ("snv", "indel", "<DUP>", "<INV>", "<INS>", "<DEL>", "<DEL>/cnv")), kinds
    np.savez_compressed(os.path.join(HERE, "g4_vcf.npz"), fasta=c.chars, snv=r.snv, snv_ave_rd=np.array(r.snv_ave_rd), ins=r.ins, del_ev=r.del_ev,
                        sv_ev=r.sv_ev, cnv=calls, vcf=np.array("".join(ref)),
                        mean=np.array([m[k] for k in ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")]))
    print("g4_vcf: %d records, %.2f MB" % (len(ref), os.path.getsize(os.path.join(HERE, "g4_vcf.npz")) / 1e6))
    shutil.rmtree(tmp)


def main_ctx():
    """g5_ctx.npz: translocation gate events of every contig of a data set with reciprocal inter-contig clusters, and the reference's
    <out>.ctx.vcf records (white-box build == prebuilt binary)."""
    from grom_b200 import hostlib
    from grom_b200.params import Params
    spec = synth.SynthSpec(contigs=[("chrA", 300_000), ("chrB", 300_000), ("chrZ", 50_000)], depth=30, seed=31, sv_classes=20, disc_frac=0.005)
    cs = synth.simulate(spec)
    tmp = tempfile.mkdtemp()
    fa, bam = synth.write_dataset(os.path.join(tmp, "g5"), cs)
    po.run_reference(bam, fa, os.path.join(tmp, "g5.vcf"), kind="ref")
    ref = [l for l in open(os.path.join(tmp, "g5.ctx.vcf")) if not l.startswith("#")]
    po.run_reference(bam, fa, os.path.join(tmp, "g5d.vcf"), kind="dist")
    assert ref == [l for l in open(os.path.join(tmp, "g5d.ctx.vcf")) if not l.startswith("#")] and len(ref) >= 4
    m = po.read_mean_file(bam)
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"], lseq=m["lseq"])
    hez, mq = po.reference_tables(20)
    out = dict(vcf=np.array("".join(ref)), names=np.array([c.name for c in cs]),
               mean=np.array([m[k] for k in ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")]))
    recs = []
    with hostlib.Bam(bam) as b:
        for tid, c in enumerate(cs):
            r = po.run_chr(prm, b.read_target(tid), c.chars, hez, mq)
            ev = r.sv_ev[r.sv_ev["cls"] >= 8]
            out[f"events_{tid}"] = ev[ev["cls"] <= 9]
            recs.append(hostlib.ctx_contig(prm, tid, r.sv_ev))
    assert hostlib.ctx_vcf(prm, [c.name for c in cs], np.concatenate(recs)).splitlines(keepends=True) == ref
    np.savez_compressed(os.path.join(HERE, "g5_ctx.npz"), **out)
    print("g5_ctx: %d records, %d candidates, %.2f MB" % (len(ref), sum(len(x) for x in recs), os.path.getsize(os.path.join(HERE, "g5_ctx.npz")) / 1e6))
    shutil.rmtree(tmp)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ctx":
        main_ctx()
    elif len(sys.argv) > 1 and sys.argv[1] == "vcf":
        main_vcf()
    elif len(sys.argv) > 1 and sys.argv[1] == "sv":
        main_sv()
    elif len(sys.argv) > 1 and sys.argv[1] == "cnv":
        main_cnv()
    else:
        main()
        main_cnv()
        main_sv()
        main_vcf()
        main_ctx()
