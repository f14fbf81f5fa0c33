"""Development aid: run the white-box reference with dump hooks on a synthetic data set and compare every dumped
per-position value with the oracle (needs oracle/_ref/GROM_ref; build container only)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from grom_b200 import hostlib
from grom_b200.params import GA_NAMES, Params
from oracle import pyoracle as po
from tools import synth

CL = ["del_f", "del_r", "dup_f", "dup_r", "inv_f1", "inv_r1", "inv_f2", "inv_r2", "ctx_f", "ctx_r"]

ap = argparse.ArgumentParser()
ap.add_argument("--seed", type=int, default=7)
ap.add_argument("--rmdup", type=int, default=0)
ap.add_argument("--sa", type=float, default=0.5)
ap.add_argument("--disc", type=float, default=0.03)
ap.add_argument("--len", type=int, default=300000)
ap.add_argument("--sv", type=float, default=10.0)
ap.add_argument("--svc", type=float, default=0.0, help="planted clusters per Mb of each further SV class")
ap.add_argument("--cnv", type=float, default=0.0)
ap.add_argument("--depth", type=float, default=30.0)
a = ap.parse_args()
spec = synth.SynthSpec(contigs=[("chrA", a.len), ("chrB", a.len // 2), ("chrZ", 50000)], depth=a.depth, seed=a.seed, dup_frac=0.05, cnv_per_mb=a.cnv, cnv_min=15000, cnv_max=30000,
                       sa_frac=a.sa, disc_frac=a.disc, sv_sites_per_mb=a.sv, munmap_frac=0.01, sv_classes=a.svc)
cs = synth.simulate(spec)
fa, bam = synth.write_dataset("/tmp/cmpref", cs)
dump = "/tmp/cmpref_dump"
os.system(f"rm -rf {dump}")
po.run_reference(bam, fa, "/tmp/cmpref.vcf", args=(["-M"] if a.rmdup else []), dump_dir=dump)
m = po.read_mean_file(bam)
print(m)
prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"], lseq=m["lseq"], rmdup=a.rmdup)
hez, mq = po.reference_tables(20)
bad_total = 0
with hostlib.Bam(bam) as b:
    for tid, c in enumerate(cs):
        batch = b.read_target(tid)
        name = c.name.lower()
        r = po.run_chr(prm, batch, c.chars, hez, mq)
        sd = po.load_scan_dump(dump, name)
        pos = sd["pos"]
        print(name, "scan", r.scan_first, r.scan_last, "dump", pos[0], pos[-1])
        for k in range(51):
            mine, ref = r.arrays[k][pos], sd["v"][:, k]
            nb = int((mine != ref).sum())
            if nb:
                j = np.nonzero(mine != ref)[0][:3]; bad_total += nb
                print(f"  {GA_NAMES[k]:16s} mismatches {nb:7d}  pos {pos[j]} mine {mine[j]} ref {ref[j]}")
        for k, cn in enumerate(CL):
            for nm, mine, ref in (("w", r.cl_w[k][pos], sd["v"][:, 51 + 3 * k]), ("rs", r.cl_rs[k][pos], sd["v"][:, 52 + 3 * k]),
                                  ("re", r.cl_re[k][pos], sd["v"][:, 53 + 3 * k]), ("dist", r.cl_dist[k][pos], sd["d"][:, k])):
                # read_start/end/dist are only meaningful where the cluster exists
                live = (r.cl_w[k][pos] != 0) | (sd["v"][:, 51 + 3 * k] != 0) if nm != "w" else np.ones(len(pos), bool)
                bad = np.nonzero((mine != ref) & live)[0]
                if bad.size:
                    bad_total += bad.size
                    print(f"  {cn}.{nm:5s} mismatches {bad.size:7d}  pos {pos[bad[:3]]} mine {mine[bad[:3]]} ref {ref[bad[:3]]}   (nonzero ref {int((sd['v'][:, 51 + 3 * k] != 0).sum())})")
        for k, cn in enumerate(("ctx_f", "ctx_r")):
            live = r.cl_w[8 + k][pos] != 0
            bad = np.nonzero((r.cl_mchr[k][pos] != sd["v"][:, 81 + k]) & live)[0]
            if bad.size:
                print(f"  {cn}.mchr mismatches {bad.size}"); bad_total += bad.size
        vcf = [l for l in open("/tmp/cmpref.vcf") if not l.startswith("#")]
        for kind, mine in (("SNV", hostlib.vcf_snv(prm, name, c.chars, r.snv, r.snv_ave_rd)), ("INS", hostlib.vcf_ins(prm, name, c.chars, r.ins)),
                           ("DEL", hostlib.vcf_smalldel(prm, name, c.chars, r.del_ev))):
            mine = po.normalise_records(mine.splitlines(keepends=True))
            if kind == "SNV":
                ref = [l for l in vcf if l.startswith(name + "\t") and l.split("\t")[2] == ""]
            elif kind == "INS":
                ref = po.normalise_records([l for l in vcf if l.startswith(name + "\t") and "\tSPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP\t" in l])
            else:
                ref = [l for l in vcf if l.startswith(name + "\t") and "\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT:SSC:ESC:HP\t" in l]
            print(f"  {kind} vcf lines mine/ref {len(mine)}/{len(ref)} identical {mine == ref}")
            if mine != ref:
                bad_total += 1
        bad = np.nonzero(r.other_len[pos] != sd["v"][:, 83])[0]
        if bad.size:
            bad_total += bad.size
            print(f"  other_len mismatches {bad.size} pos {pos[bad[:3]]} mine {r.other_len[pos][bad[:3]]} ref {sd['v'][:, 83][bad[:3]]}  (max ref {sd['v'][:, 83].max()})")
        # structural-variant candidate lists at the end of the scan: product host stage on the oracle's gate events vs the reference's lists
        ref_l = po.load_svlist_dump(dump, name)
        mine_l = po.normalise_sv_lists(hostlib.sv_lists(prm, r.sv_ev))
        for k in ("dup", "del", "inv_f", "inv_r", "ins", "ctx_f", "ctx_r"):
            same = len(ref_l[k]) == len(mine_l[k]) and ref_l[k].tobytes() == mine_l[k].tobytes()
            print(f"  sv list {k:6s} ref {len(ref_l[k]):5d} mine {len(mine_l[k]):5d} identical {same}")
            if not same:
                bad_total += 1
                n = min(len(ref_l[k]), len(mine_l[k]))
                for i in range(n):
                    if ref_l[k][i].tobytes() != mine_l[k][i].tobytes():
                        print("    first difference at", i, "\n     ref ", ref_l[k][i], "\n     mine", mine_l[k][i]); break
        # the complete record text of the contig, in the reference's order
        cn = po.cnv_run(prm, name, c.chars, r["gc"], r["acgt"], r["rd_mq"], r["rd_rd"], r["rd_low"], seed=1)
        from grom_b200.params import CNV_CALL_DTYPE
        calls = np.zeros(len(cn.dels) + len(cn.dups), dtype=CNV_CALL_DTYPE)
        for k, src in enumerate((cn.dels, cn.dups)):
            sl = slice(0, len(cn.dels)) if k == 0 else slice(len(cn.dels), None)
            calls["start"][sl] = src["start"]; calls["end"][sl] = src["end"]; calls["kind"][sl] = k; calls["z"][sl] = src["z"]
            calls["pvalue"][sl] = src["p"]; calls["cn"][sl] = src["cn"]; calls["cn_sd"][sl] = src["cs"]
        mine_all = po.normalise_records(hostlib.vcf_contig(prm, name, c.chars, r.snv, r.snv_ave_rd, r.ins, r.del_ev, r.sv_ev, calls).splitlines(keepends=True))
        ref_all = po.normalise_records([l for l in vcf if l.startswith(name + "\t")])
        same = mine_all == ref_all
        kinds = {}
        for l in ref_all:
            t = l.split("\t")[4]; kinds[t if t.startswith("<") else "seq"] = kinds.get(t if t.startswith("<") else "seq", 0) + 1
        print(f"  FULL contig records: ref {len(ref_all)} mine {len(mine_all)} identical {same}   {kinds}")
        if not same:
            bad_total += 1
            import difflib
            for l in list(difflib.unified_diff(ref_all, mine_all, lineterm="", n=0))[:12]:
                print("     ", l.rstrip()[:200])
print("TOTAL MISMATCHES", bad_total)
allrec = [l for l in open("/tmp/cmpref.vcf") if not l.startswith("#")]
print("reference records:", len(allrec), "of which SV/CNV classes not yet produced:", sum(1 for l in allrec if l.split("\t")[4].startswith("<") and "SSC:ESC" not in l and "SSC:HP" not in l))
