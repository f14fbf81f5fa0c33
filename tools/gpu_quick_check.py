"""A few seconds on a GPU box, no pytest, no torch: the Python driver (grom_b200.pipeline) on the committed golden data set, compact FASTA
through the C reader and the gzip one through the Python reader, three lanes and one, whole contigs and contigs in pieces; every record against the reference's VCF
(tests/golden/g1_*.npz).  Prints one line per case; exit code 1 on any difference."""
import gzip
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grom_b200 import pipeline  # noqa: E402
from grom_b200.params import Params  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (checker only)

G = os.path.join(ROOT, "tests", "golden")
plain = "/tmp/g1_quick.fa"
open(plain, "wb").write(gzip.open(os.path.join(G, "g1.fa.gz"), "rb").read())
bad = 0
for tag, rmdup in (("default", 0), ("rmdup", 1)):
    g = np.load(os.path.join(G, f"g1_{tag}.npz"))
    ref = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#")]
    for fa, lanes in ((plain, 3), (os.path.join(G, "g1.fa.gz"), 1)):
        t0 = time.time()
        text, prm = pipeline.call_variants(os.path.join(G, "g1.bam"), fa, Params.default(rmdup=rmdup), lanes=lanes)
        mine = "".join(text[t] for t in sorted(text)).splitlines(keepends=True)
        ok = po.normalise_records(mine) == po.normalise_records(ref)
        m = g["mean"]
        ok_stats = (prm.insert_mean, prm.lseq, prm.insert_min, prm.insert_max) == (int(max(m[0], m[1])), int(m[1]), int(m[2]), int(m[3]))
        print(f"{tag:8s} lanes={lanes} fasta={'gz' if fa.endswith('.gz') else 'plain'}: {len(mine)} records, equal to the reference: {ok}, statistics: {ok_stats}, {time.time() - t0:.2f} s", flush=True)
        bad += (not ok) + (not ok_stats)
# the same contigs decoded and pushed in pieces (gromhost_bam_iter_* -> consecutive gromgpu_push_reads): small decode windows so that the
# pieces of this small data set are many
os.environ["GROMHOST_WINDOW_BLOCKS"] = "2"
for tag, rmdup in (("default", 0), ("rmdup", 1)):
    g = np.load(os.path.join(G, f"g1_{tag}.npz"))
    ref = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#")]
    t0 = time.time()
    text, prm = pipeline.call_variants(os.path.join(G, "g1.bam"), plain, Params.default(rmdup=rmdup), lanes=3, slice_reads=500)
    mine = "".join(text[t] for t in sorted(text)).splitlines(keepends=True)
    ok = po.normalise_records(mine) == po.normalise_records(ref)
    print(f"{tag:8s} lanes=3 in pieces of >= 500 reads: {len(mine)} records, equal to the reference: {ok}, {time.time() - t0:.2f} s", flush=True)
    bad += not ok
sys.exit(1 if bad else 0)
