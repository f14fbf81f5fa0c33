"""Structural-variant scan, CPU side: the oracle's gate events and the product host list builder (grom_b200/host/svlists.c) against the
reference's own candidate lists at the end of the per-position scan (tests/golden/g3_svlists.npz, dumped by the white-box build at
src/GROM.c:15164; made by tests/golden/make_golden.py sv)."""
import os

import numpy as np

from util import GOLDEN, golden_batches, golden_params, load_golden_fasta, tables_7digit
from grom_b200 import hostlib
from grom_b200.params import Params, SV_EVENT_DTYPE
from oracle import pyoracle as po

LISTS = ("dup", "del", "inv_f", "inv_r", "ins", "ctx_f", "ctx_r")


def g3():
    return np.load(os.path.join(GOLDEN, "g3_svlists.npz"))


def test_list_builder_reproduces_reference_lists_all_classes():
    g = g3()
    m = g["g3_mean"]
    prm = Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]))
    total = 0
    assert set(g["g3_chra_events"]["cls"].tolist()) | set(g["g3_chrb_events"]["cls"].tolist()) == set(range(12))
    for chrom in ("chra", "chrb"):
        ev = g[f"g3_{chrom}_events"]
        mine = po.normalise_sv_lists(hostlib.sv_lists(prm, ev))
        for k in LISTS:
            ref = g[f"g3_{chrom}_{k}"]
            assert len(ref) > 0 and mine[k].tobytes() == ref.tobytes(), (chrom, k)
            total += len(ref)
        # the builder sorts into scan order itself: any permutation of the events gives the same lists
        rng = np.random.default_rng(1)
        again = po.normalise_sv_lists(hostlib.sv_lists(prm, ev[rng.permutation(len(ev))]))
        assert all(again[k].tobytes() == mine[k].tobytes() for k in LISTS)
    assert total > 15000
    empty = hostlib.sv_lists(prm, np.zeros(0, dtype=SV_EVENT_DTYPE))
    assert all(len(v) == 0 for v in empty.values())


def test_oracle_gates_and_list_builder_on_golden_bam():
    """g1 (committed BAM): oracle gate events -> host list builder == the reference's lists for every contig."""
    g = g3()
    names, batches = golden_batches()
    fasta = load_golden_fasta()
    hez, mq = tables_7digit()
    d = np.load(os.path.join(GOLDEN, "g1_default.npz"))
    prm = golden_params(d, 0)
    n = 0
    for tid, name in enumerate(names):
        r = po.run_chr(prm, batches[tid], fasta[name], hez, mq)
        mine = po.normalise_sv_lists(hostlib.sv_lists(prm, r.sv_ev))
        for k in LISTS:
            ref = g[f"g1_{name.lower()}_{k}"]
            assert mine[k].tobytes() == ref.tobytes(), (name, k, len(mine[k]), len(ref))
            n += len(ref)
    assert n >= 0          # the small random g1 contigs have few (or no) passing gates; the rich case is g3 above


def test_translocation_records_against_reference_ctx_vcf():
    """tests/golden/g5_ctx.npz: per-contig translocation gate events -> candidate merge + filter (gromhost_ctx_contig) -> mate pairing
    across contigs and BND records (gromhost_ctx_vcf) == the reference's <out>.ctx.vcf."""
    g = np.load(os.path.join(GOLDEN, "g5_ctx.npz"))
    m = g["mean"]
    prm = Params.default(insert_mean=int(max(m[0], m[1])), lseq=int(m[1]), insert_min=int(m[2]), insert_max=int(m[3]))
    names = [str(x) for x in g["names"]]
    recs = [hostlib.ctx_contig(prm, tid, g[f"events_{tid}"]) for tid in range(len(names))]
    text = hostlib.ctx_vcf(prm, names, np.concatenate(recs))
    assert text == str(g["vcf"]) and text.count("SVTYPE=BND") >= 4
    ids = [int(l.split("\t")[2]) for l in text.splitlines()]
    mates = [int(l.split("MATEID=")[1].split("\t")[0]) for l in text.splitlines()]
    assert sorted(ids) == sorted(mates)                       # every kept record's mate is kept too
    assert hostlib.ctx_vcf(prm, names, np.concatenate(recs)[:0]) == ""
