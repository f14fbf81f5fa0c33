"""Read-depth CNV path on the GPU (gromgpu_chr_cnv) against the CPU oracle (oracle/grom_oracle_cnv.c, pinned against the
reference's own dumps): pre-statistics, per-bin distributions, mask and z list, window-length sd table, calls, copy number,
p-values and the VCF text of the host writer.  Doubles are compared bit for bit unless noted."""
import numpy as np
import pytest

from util import tables_7digit
from grom_b200 import gpu, hostlib
from grom_b200.params import Params
from oracle import pyoracle as po
from tools import synth

pytestmark = pytest.mark.gpu


def _dataset(seed, length, depth, cnv_per_mb, **kw):
    spec = synth.SynthSpec(contigs=[("chrA", length)], depth=depth, seed=seed, cnv_per_mb=cnv_per_mb, disc_frac=0.005, sv_sites_per_mb=1.0,
                           low_mapq_frac=0.05, **kw)
    return synth.simulate(spec)[0]


def _run_both(prm, c, ploidy=None):
    hez, mq = tables_7digit()
    gpu.init(0, hez, mq, prm)
    ref = po.run_chr(prm, c.batch, c.chars, hez, mq)
    gc, acgt = ref["gc"], ref["acgt"]
    o = po.cnv_run(prm, "chra", c.chars, gc, acgt, ref["rd_mq"], ref["rd_rd"], ref["rd_low"], ploidy=ploidy, seed=prm.rand_seed,
                   sample_cap=prm.sample_lists_len, min_win=prm.min_rd_window_len, max_win=prm.max_rd_window_len)
    with gpu.Chromosome(0, c.chars) as ch:
        ch.push_reads(c.batch)
        ch.run()
        g = ch.cnv(ploidy=ploidy)
        z, mask, mqm = ch.cnv_fetch("z"), ch.cnv_fetch("mask"), ch.cnv_fetch("mq_mean")
        # the CNV stage must leave the depth arrays of the run untouched
        assert np.array_equal(ch.fetch("rd_mq"), ref["rd_mq"])
    return o, g, z, mask, mqm


def _check(o, g, z, mask, mqm, prm):
    assert g.chr_ave == o.chr_ave and g.blk_ave == o.blk_ave and g.biased_repeat == o.biased
    assert abs(g.chr_sd - o.chr_sd) <= 1e-9 * max(1.0, abs(o.chr_sd))       # summed per depth value, not per position
    assert g.n_repeats == len(o.repeats) and g.n_sample_blocks == len(o.sample_blocks)
    assert np.array_equal(g.n, o.windows)
    for a, b, nm in ((g.ave, o.ave, "ave"), (g.sd, o.sd, "sd"), (g.del_thr, o.del_thr, "del_thr"), (g.dup_thr, o.dup_thr, "dup_thr")):
        assert np.array_equal(a, b), nm
    assert np.array_equal(mqm, np.minimum(o.mq_mean, 255).astype(np.uint8))
    assert np.array_equal(mask, o.mask), f"mask differs at {np.nonzero(mask != o.mask)[0][:5]}"
    assert np.array_equal(z, o.z), f"z differs at {np.nonzero(z != o.z)[0][:5]}"
    assert np.array_equal(g.win_cnt, o.win_cnt)
    assert np.array_equal(g.win_sd, o.win_sd)
    dels, dups = g.calls[g.calls["kind"] == 0], g.calls[g.calls["kind"] == 1]
    for mine, ref in ((dels, o.dels), (dups, o.dups)):
        assert np.array_equal(mine["start"], ref["start"]) and np.array_equal(mine["end"], ref["end"])
        for f, rf in (("z", "z"), ("cn", "cn"), ("cn_sd", "cs"), ("pvalue", "p")):
            assert np.array_equal(mine[f], ref[rf]), f
    assert hostlib.vcf_cnv(prm, "chra", g.calls) == o.vcf


def test_cnv_matches_oracle_default():
    c = _dataset(seed=5, length=3_000_000, depth=30, cnv_per_mb=0.7)
    prm = Params.default()
    o, g, z, mask, mqm = _run_both(prm, c)
    assert len(o.dels) > 0 and len(o.dups) > 0 and (mask == 0).sum() > 1_000_000
    assert "<DEL>" in o.vcf                     # at least one planted loss survives the -V filter
    _check(o, g, z, mask, mqm, prm)


def test_cnv_matches_oracle_tetraploid_A4():
    c = _dataset(seed=8, length=900_000, depth=40, cnv_per_mb=2.3)
    prm = Params.default(ploidy=4, windows_sampling_factor=4)
    o, g, z, mask, mqm = _run_both(prm, c)
    _check(o, g, z, mask, mqm, prm)


def test_cnv_reservoir_and_small_windows():
    """A small list capacity forces the libc-rand reservoir path (src/GROM.c:18391-18398); short windows exercise the frame logic."""
    c = _dataset(seed=9, length=700_000, depth=15, cnv_per_mb=3.0)
    prm = Params.default(sample_lists_len=150, rand_seed=7, min_rd_window_len=50, max_rd_window_len=3000, windows_sampling_factor=3)
    o, g, z, mask, mqm = _run_both(prm, c)
    assert int(o.windows.max()) == 150
    _check(o, g, z, mask, mqm, prm)


def test_cnv_biased_repeat_override():
    """> 100 thinned (AT)n runs make AT the most biased repeat type: sampled repeat lists and the z override around every run."""
    c = _dataset(seed=21, length=1_200_000, depth=30, cnv_per_mb=0.9, at_repeats=260, cnv_min=15000, cnv_max=30000)
    prm = Params.default()
    o, g, z, mask, mqm = _run_both(prm, c)
    assert o.biased == 3 and "<DEL>" in o.vcf
    _check(o, g, z, mask, mqm, prm)


def test_cnv_host_scan_fallback_gives_the_same_calls(monkeypatch):
    """The segmentation has two routes: hop on the device over the successor table, or (seed tables too large / too many unresolved
    seeds) the host scan in parallel pieces over the packed records.  Both must agree."""
    c = _dataset(seed=5, length=3_000_000, depth=30, cnv_per_mb=0.7)
    prm = Params.default()
    hez, mq = tables_7digit()
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(0, c.chars) as ch:
        ch.push_reads(c.batch)
        ch.run()
        a = ch.cnv()
        monkeypatch.setenv("GROMGPU_CNV_HOST_SCAN", "1")             # host walk over the device-evaluated seed tables
        b = ch.cnv()
        monkeypatch.delenv("GROMGPU_CNV_HOST_SCAN")
        monkeypatch.setenv("GROMGPU_CNV_NO_SEED_TABLES", "1")        # every seed evaluated on the host
        d = ch.cnv()
        monkeypatch.delenv("GROMGPU_CNV_NO_SEED_TABLES")
    assert len(a.calls) > 20 and a.calls.tobytes() == b.calls.tobytes() == d.calls.tobytes()
    assert np.array_equal(a.win_sd, b.win_sd)


def test_cnv_short_contig_is_empty():
    """A contig shorter than the GC window has no analysed span (lo >= hi): no calls, no error."""
    rng = np.random.default_rng(3)
    chars = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 1000)].copy()
    batch = synth.batch_from_records(0, [dict(pos=100 + 50 * i, cigar=[(0, 100)], seq="A" * 100, qual=30, flag=0, mapq=60) for i in range(5)])
    prm = Params.default()
    hez, mq = tables_7digit()
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(0, chars) as ch:
        ch.push_reads(batch)
        ch.run()
        g = ch.cnv()
    assert len(g.calls) == 0 and g.n_samples == 0


def test_cnv_size_independent_properties_large():
    """8 Mb / 30x with planted copy-number segments: properties that hold at any size -- a second run is identical, calls are sorted and
    disjoint per kind, every window length up to the frame size has observations, the mask covers exactly what it must, the planted
    one-copy losses are recovered -- plus errors of the C ABI."""
    spec = synth.SynthSpec(contigs=[("chrL", 8_000_000)], depth=30, seed=31, simple=True, cnv_per_mb=0.5, cnv_min=20000, cnv_max=60000)
    c = synth.simulate(spec)[0]
    prm = Params.default()
    hez, mq = tables_7digit()
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(0, c.chars) as ch:
        with pytest.raises(gpu.GromGpuError, match="chr_run first"):
            ch.cnv()
        ch.push_reads(c.batch)
        ch.run()
        g1 = ch.cnv(); z1 = ch.cnv_fetch("z"); m1 = ch.cnv_fetch("mask"); d1 = ch.cnv_fetch("depth")
        g2 = ch.cnv(); z2 = ch.cnv_fetch("z")
        with pytest.raises(gpu.GromGpuError, match="bad range"):
            ch.cnv_fetch("z", 10, len(c.chars) + 5)
        rd = ch.fetch("rd_rd") + ch.fetch("rd_low")
        acgt = ch.fetch("acgt")
    assert g1.calls.tobytes() == g2.calls.tobytes() and np.array_equal(z1, z2) and np.array_equal(g1.win_sd, g2.win_sd)
    assert np.array_equal(d1, rd)
    M = prm.insert_mean
    lo, hi = M - 1, len(c.chars) - (2 * M - 1)
    assert m1[:lo].all() and m1[hi:].all() and np.all(m1[lo:hi][acgt[lo:hi] < 99] == 1) and (m1 == 0).mean() > 0.9
    assert np.all(z1[m1 == 1] == 0)
    assert np.all(g1.win_cnt[prm.min_rd_window_len:] > 1) and np.all(g1.win_sd[prm.min_rd_window_len:] > 0)
    # longer windows average more positions: the null sd shrinks (monotone up to sampling noise)
    assert g1.win_sd[100] > g1.win_sd[1000] > g1.win_sd[10000]
    for kind in (0, 1):
        k = g1.calls[g1.calls["kind"] == kind]
        assert np.all(np.diff(k["start"]) > 0) and np.all(k["start"][1:] > k["end"][:-1]) and np.all(k["end"] >= k["start"])
        assert np.all((k["z"] >= 3) & (k["pvalue"] < 0.5))        # the reference's erf variant can push strong calls to 0 or slightly below
    dels = g1.calls[g1.calls["kind"] == 0]
    for a, b, cn in c.truth["cnv"]:
        if cn == 1:      # one-copy loss: most of the segment is inside deletion calls
            cov = sum(max(0, min(b, int(e)) - max(a, int(s))) for s, e in zip(dels["start"], dels["end"]))
            assert cov > 0.8 * (b - a), (a, b, cov)
