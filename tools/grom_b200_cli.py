"""Command-line front end with the reference's flag letters (src/GROM.c:21907-22106) over grom_b200.pipeline:
    python tools/grom_b200_cli.py -i in.bam -r ref.fa -o out.vcf [-M] [-q 20] [-b 20] [-v 0.001] [-e 1e-10] [-V 1e-9] [-p 2] [-g 0] [-A 2]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grom_b200 import pipeline
from grom_b200.params import Params

ap = argparse.ArgumentParser()
ap.add_argument("-i", required=True); ap.add_argument("-r", required=True); ap.add_argument("-o", required=True)
ap.add_argument("-M", action="store_true"); ap.add_argument("-q", type=int, default=20); ap.add_argument("-b", type=int, default=20)
ap.add_argument("-v", type=float, default=0.001); ap.add_argument("-e", type=float, default=1e-10); ap.add_argument("-V", type=float, default=1e-9)
ap.add_argument("-p", type=int, default=2); ap.add_argument("-g", type=int, default=0); ap.add_argument("-A", type=int, default=2)
ap.add_argument("--device", type=int, default=0)
a = ap.parse_args()
prm = Params.default(rmdup=int(a.M), min_mapq=a.q, rd_min_mapq=a.q, min_base_qual=a.b, pval_threshold=a.v, pval_threshold1=a.v, pval_insertion=a.e,
                     rd_pval_threshold=a.V, ploidy=a.p, gender=a.g, windows_sampling_factor=a.A)
from grom_b200 import hostlib
ctx = {}
text, prm = pipeline.call_variants(a.i, a.r, prm, device=a.device, ctx_out=ctx)
pipeline.write_vcf(a.o, text, a.r)
with hostlib.Bam(a.i) as b:
    names = b.names
ctx_path = a.o[:-4] + ".ctx.vcf" if a.o.endswith(".vcf") else a.o + ".ctx"          # the reference's naming, src/GROM.c:22431-22445
pipeline.write_ctx_vcf(ctx_path, pipeline.ctx_vcf_text(prm, names, ctx), a.r)
print(f"insert_mean {prm.insert_mean} lseq {prm.lseq} insert_min {prm.insert_min} insert_max {prm.insert_max}; "
      f"{sum(t.count(chr(10)) for t in text.values())} records over {len(text)} contigs -> {a.o}")
