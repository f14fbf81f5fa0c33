"""Write a synthetic BAM + FASTA sample of the bench workload to a directory (child process of bench.py's reference arm, so that the
process that times the reference maps none of this repository's libraries).  Prints one JSON line: paths, contigs, aligned bases."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--contigs", type=int, default=8)
ap.add_argument("--mb", type=float, default=1.5)
ap.add_argument("--depth", type=float, default=30.0)
ap.add_argument("--seed", type=int, default=4242)
ap.add_argument("--dummy", type=int, default=60_000, help="length of the last contig (-P >= 2 silently drops the last FASTA contig, reference src/GROM.c:20999)")
a = ap.parse_args()
spec = workloads.chr20_spec(mb=a.mb, depth=a.depth, seed=a.seed, names=True, dummy_len=a.dummy)
spec.contigs = [(f"chr{i + 1}", int(a.mb * 1e6)) for i in range(a.contigs)] + [("chrzz", a.dummy)]
cs = synth.simulate(spec)
os.makedirs(a.out, exist_ok=True)
fa, bam = synth.write_dataset(os.path.join(a.out, "sample"), cs)
bases = sum(c.batch.aligned_bases() for c in cs[:-1])
print(json.dumps({"fasta": fa, "bam": bam, "contigs": a.contigs, "mb": a.mb, "aligned_bases": int(bases), "reads": int(sum(c.batch.n_reads for c in cs[:-1]))}))
