"""Tilapia stand-in (BASELINE.json configs 1 and 2): the reference ships `test_data/oreNil2_GL831235-1.fa` and a golden VCF, but the
BAM it was made from is a missing blob.  This script simulates 5x paired reads (tools/synth.py, seed 1) ON THAT REAL FASTA -- soft-masked
lower case, 153 k N -- runs the reference itself on them (oracle/_ref/GROM_dist, its prebuilt binary; default flags and -M) and commits

    tests/golden/oreNil2_GL831235-1.fa.gz       the reference's test FASTA (data, not source), gzip -9
    tests/golden/g6_tilapia_standin.npz         the reference's VCF records (default, -M), .ctx.vcf bodies, library statistics, and a
                                                digest of the simulated reads (guards the RNG stream the tests regenerate them with)

Run in the build container:  python tests/golden/make_tilapia_standin.py
"""
import gzip
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tools import synth  # noqa: E402

SRC = "/root/reference/test_data/oreNil2_GL831235-1.fa"
NAME = "GL831235-1"


def load_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    assert lines[0].startswith(b">" + NAME.encode())
    return np.frombuffer(b"".join(l.strip() for l in lines[1:]), dtype=np.uint8).copy()


def standin_spec(length):
    return synth.SynthSpec(contigs=[(NAME, length), ("chrzz", 30_000)], depth=5, seed=1, dup_frac=0.05, disc_frac=0.02, sa_frac=0.8,
                           sv_sites_per_mb=6.0, sv_classes=3, cnv_per_mb=1.5, cnv_min=15_000, cnv_max=40_000)


def simulate(chars):
    return synth.simulate(standin_spec(len(chars)), references={NAME: chars})


def reads_digest(cs):
    h = hashlib.sha1()
    for c in cs:
        b = c.batch
        for a in (b.pos, b.flag, b.mapq, b.tlen, b.cigar, b.seq4, b.qual):
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    chars = load_fasta(SRC)
    with open(SRC, "rb") as f, gzip.open(os.path.join(HERE, "oreNil2_GL831235-1.fa.gz"), "wb", 9) as g:
        g.write(f.read())
    cs = simulate(chars)
    tmp = tempfile.mkdtemp()
    fa, bam = synth.write_dataset(os.path.join(tmp, "til"), cs)
    out = dict(digest=np.array(reads_digest(cs)), n_reads=np.array([c.batch.n_reads for c in cs]))
    for tag, args in (("default", []), ("rmdup", ["-M"])):
        vcf = os.path.join(tmp, tag + ".vcf")
        po.run_reference(bam, fa, vcf, args=args, kind="dist")
        out[f"vcf_{tag}"] = np.array("".join(l for l in open(vcf) if not l.startswith("#")))
        out[f"ctx_{tag}"] = np.array("".join(l for l in open(os.path.join(tmp, tag + ".ctx.vcf")) if not l.startswith("#")))
        if tag == "default":
            out["mean"] = np.array([po.read_mean_file(bam)[k] for k in ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")])
        print(tag, out[f"vcf_{tag}"].item().count("\n"), "records,", out[f"ctx_{tag}"].item().count("\n"), "translocation records")
    np.savez_compressed(os.path.join(HERE, "g6_tilapia_standin.npz"), **out)


if __name__ == "__main__":
    main()
