"""The named synthetic workloads of BASELINE.json / SURVEY.md 8(d), shared by bench.py, the probes and the tests.

config 3 (`chr20_spec`): one chr20-sized contig at 30x, 2x150 bp, insert N(400,40), with every evidence class the hot path
handles: 3 % low-MAPQ reads, 0.2 % substitution errors, planted SNVs every ~1 kb (hom/het), small indels of 1-20 bp every
~10 kb, 2 % soft-clipped reads (SA tags on most clips >= 20 bp), hard clips, reference skips, 1 % discordant pairs of every
orientation class, mate-unmapped and inter-contig pairs, per 10 Mb 5 clustered deletions + 2 tandem duplications + 2+2
inversion breakpoints + 2 translocations + 2 insertions (supporting pairs and split reads), copy-number segments of
20-200 kb at ploidy +-1 (every fourth loss a full loss), 5 % exact PCR duplicates for -M.  A small second contig receives the
inter-contig mates (and keeps a `-P >= 2` reference run from silently dropping the real contig, reference src/GROM.c:20999).

Nothing here is on the product path.
"""
from __future__ import annotations

from tools import synth


def chr20_spec(mb: float = 64.0, depth: float = 30.0, seed: int = 20, name: str = "chr20", names: bool = False, cnv_per_mb: float = 0.25,
               dummy_len: int = 1_000_000) -> synth.SynthSpec:
    return synth.SynthSpec(contigs=[(name, int(mb * 1e6)), ("chrzz", dummy_len)], depth=depth, seed=seed, simple=False, names=names,
                           dup_frac=0.05, disc_frac=0.01, clip_frac=0.02, sa_frac=0.8, low_mapq_frac=0.03, indel_every=10_000,
                           sv_sites_per_mb=0.5, sv_classes=0.2, cnv_per_mb=cnv_per_mb, cnv_min=20_000, cnv_max=200_000)


def tetraploid_spec(mb: float = 64.0, depth: float = 100.0, seed: int = 100, name: str = "chr1", names: bool = False) -> synth.SynthSpec:
    """config 5: 100x, run with -p 4 -A 4 (the generator is diploid; the flags are what stress the window sweep)."""
    s = chr20_spec(mb, depth, seed, name, names)
    return s


def params_config3(rmdup: int = 1, **kw):
    from grom_b200.params import Params
    return Params.default(insert_mean=400, insert_min=170, insert_max=520, lseq=150, rmdup=rmdup, **kw)
