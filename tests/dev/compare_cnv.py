"""Development aid: run the white-box reference with dump hooks on a synthetic data set with planted copy-number segments and
compare the CNV path state (pre-statistics, distributions, z list, window sd table, calls, VCF records) with the oracle."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from grom_b200 import hostlib
from grom_b200.params import GA, Params
from oracle import pyoracle as po
from tools import synth

ap = argparse.ArgumentParser()
ap.add_argument("--seed", type=int, default=3)
ap.add_argument("--len", type=int, default=2_500_000)
ap.add_argument("--depth", type=float, default=20)
ap.add_argument("--cnv", type=float, default=4.0)
ap.add_argument("--A", type=int, default=2)
ap.add_argument("--ploidy", type=int, default=2)
ap.add_argument("--at", type=int, default=0, help="planted thinned (AT)n runs")
ap.add_argument("--cnvmax", type=int, default=120000)
a = ap.parse_args()
spec = synth.SynthSpec(contigs=[("chrA", a.len), ("chrx", a.len // 2), ("chrZ", 50000)], depth=a.depth, seed=a.seed, cnv_per_mb=a.cnv,
                       disc_frac=0.005, sv_sites_per_mb=1.0, low_mapq_frac=0.05, at_repeats=a.at, cnv_min=min(20000, a.cnvmax // 2), cnv_max=a.cnvmax)
cs = synth.simulate(spec)
fa, bam = synth.write_dataset("/tmp/cmpcnv", cs)
dump = "/tmp/cmpcnv_dump"
os.system(f"rm -rf {dump}")
t0 = time.time()
args = ["-A", a.A, "-p", a.ploidy] + (["-g", 1] if a.ploidy else [])
po.run_reference(bam, fa, "/tmp/cmpcnv.vcf", args=args, dump_dir=dump, seed=5)
print("reference run %.1fs" % (time.time() - t0))
m = po.read_mean_file(bam)
prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"], lseq=m["lseq"],
                     windows_sampling_factor=a.A, ploidy=a.ploidy, gender=1)
vcf = [l for l in open("/tmp/cmpcnv.vcf") if not l.startswith("#")]
bad_total = 0


def cmp(name, mine, ref):
    global bad_total
    mine, ref = np.asarray(mine), np.asarray(ref)
    if mine.shape != ref.shape:
        print(f"  {name}: shape {mine.shape} vs {ref.shape}"); bad_total += 1; return
    if mine.dtype.kind == "f":
        bad = ~((mine == ref) | (np.isnan(mine) & np.isnan(ref)))
    else:
        bad = mine != ref
    if bad.any():
        j = np.nonzero(bad.reshape(-1))[0][:4]
        print(f"  {name}: {int(bad.sum())} mismatches at {j} mine {mine.reshape(-1)[j]} ref {ref.reshape(-1)[j]}"); bad_total += int(bad.sum())


for tid, c in enumerate(cs):
    name = c.name.lower()
    if not os.path.exists(os.path.join(dump, f"cnv_{name}.bin")):
        print(name, "no cnv dump"); continue
    dep = po.load_depth_dump(dump, name)
    gcd = po.load_gc_dump(dump, name)
    pl = a.ploidy       # the -g 1 halving on chrx (src/GROM.c:17024) tests a name buffer that is only filled in tumour mode: dead code
    t0 = time.time()
    r = po.cnv_run(prm, name, c.chars, gcd[0], gcd[1], dep[0], dep[1], dep[2], ploidy=pl, seed=5)
    print(name, "oracle cnv %.1fs" % (time.time() - t0), "dels", len(r.dels), "dups", len(r.dups), "biased", r.biased, "sample blocks", r.sample_blocks.tolist()[:3])
    pre = po.load_cnvpre_dump(dump, name)
    d = po.load_cnv_dump(dump, name)
    for k in ("nblocks", "repeats", "chr_ave", "chr_sd", "rep_ave", "rep_sd", "rep_cnt", "biased", "blk_ave", "sample_blocks"):
        cmp("pre." + k, getattr(r, k), pre[k])
    for k in ("z", "mask", "win_sd", "win_cnt", "ave", "sd", "del_thr", "dup_thr", "windows", "n_high", "n_low"):
        cmp(k, getattr(r, k), d[k])
    for k, mine in (("dels", r.dels), ("dups", r.dups)):
        for f in ("start", "end", "z", "cn", "cs"):
            cmp(f"{k}.{f}", mine[f], d[k][f])
    ref_lines = [l for l in vcf if l.startswith(name + "\t") and "\tSD:Z:CN:CS\t" in l]
    mine = r.vcf.splitlines(keepends=True)
    print(f"  CNV VCF records: ref {len(ref_lines)} mine {len(mine)} identical {mine == ref_lines}")
    bad_total += mine != ref_lines
    print("  truth:", sorted(c.truth.get("cnv", []))[:10]); print("  dels:", [(int(x["start"]), int(x["end"]), round(float(x["z"]), 2), round(float(x["cn"]), 2), float(x["p"])) for x in r.dels][:12])
print("TOTAL MISMATCHES", bad_total)
