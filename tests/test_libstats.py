"""Library statistics (find_insert_mean, src/GROM.c:1205-1318; grom_b200/host/libstats.c): insert mean / min / max, read length and the
mapped-read figure decide every window size and gate downstream, so they are pinned on the reference's own `<bam>.mean` cache file
(src/GROM.c:994-1026): the committed data set always, freshly generated libraries where oracle/_ref/GROM_ref exists."""
import os

import numpy as np
import pytest

from util import GOLDEN, golden_batches
from grom_b200 import hostlib
from oracle import pyoracle as po
from tools import synth

KEYS = ("insert_mean", "lseq", "insert_min", "insert_max", "mapped_reads")


def test_golden_bam_statistics_equal_the_reference_mean_file():
    _, batches = golden_batches()
    st = hostlib.library_stats(batches, 20)
    m = np.load(os.path.join(GOLDEN, "g1_default.npz"))["mean"]
    assert [st[k] for k in KEYS] == [int(x) for x in m]


def test_no_reads_is_an_error():
    _, batches = golden_batches()
    with pytest.raises(RuntimeError, match="no reads"):
        hostlib.library_stats([], 20)


@pytest.mark.skipif(not po.have_reference("ref"), reason="oracle/_ref/GROM_ref not built")
@pytest.mark.parametrize("seed,read_len,ins_mean,ins_sd,q", [(61, 150, 400.0, 40, 20), (62, 100, 250.0, 15, 20), (63, 75, 500.0, 90, 30),
                                                             (64, 250, 420.0, 60, 4)])
def test_statistics_equal_the_live_reference(tmp_path, seed, read_len, ins_mean, ins_sd, q):
    spec = synth.SynthSpec(contigs=[("chrQ", 60_000), ("chrR", 30_000), ("chrZ", 10_000)], depth=15, seed=seed, read_len=read_len,
                           ins_mean=ins_mean, ins_sd=ins_sd, ins_floor=read_len + 20, dup_frac=0.02, clip_frac=0.04, disc_frac=0.03,
                           sa_frac=0.5, munmap_frac=0.01, low_mapq_frac=0.05)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), cs)
    po.run_reference(bam, fa, str(tmp_path / "o.vcf"), args=(["-q", q] if q != 20 else []))
    ref = po.read_mean_file(bam)
    with hostlib.Bam(bam) as b:
        st = hostlib.library_stats([b.read_target(t) for t in range(len(b.names))], q)
        st_file = b.library_stats(q)               # the pass the drivers run: straight over the file
    assert {k: st[k] for k in KEYS} == {k: ref[k] for k in KEYS}
    assert st_file == st


def test_statistics_straight_from_the_file_equal_the_per_contig_feed(tmp_path, monkeypatch):
    """gromhost_bam_library_stats (windowed pass over the blocks, core fields only) against the accumulator fed with whole batches: the
    committed BAM, and a file whose records straddle many windows (one thread -> windows of 32 blocks), with and without an index."""
    from util import GOLDEN
    names, batches = golden_batches()
    want = hostlib.library_stats(batches, 20)
    with hostlib.Bam(os.path.join(GOLDEN, "g1.bam")) as b:
        assert b.library_stats(20) == want and b.library_stats(20, threads=1) == want and b.library_stats(20, threads=3) == want
    spec = synth.SynthSpec(contigs=[("c1", 400_000), ("c2", 90_000)], depth=25, seed=71, dup_frac=0.03, clip_frac=0.04, disc_frac=0.03, munmap_frac=0.01,
                           unpaired_frac=0.01, low_mapq_frac=0.05)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "s"), cs)
    with hostlib.Bam(bam) as b:
        want = hostlib.library_stats([b.read_target(t) for t in range(2)], 30)
        assert os.path.getsize(bam) > 3 * 32 * 20_000                       # several one-thread windows
        for thr in (1, 2, 5):
            assert b.library_stats(30, threads=thr) == want
    os.remove(bam + ".bai")
    with hostlib.Bam(bam) as b:
        assert b.library_stats(30, threads=1) == want


def test_statistics_of_a_file_without_usable_reads(tmp_path):
    spec = synth.SynthSpec(contigs=[("c1", 20_000)], depth=3, seed=72)
    cs = synth.simulate(spec)
    cs[0].batch.flag[:] |= 4                                                # everything unmapped
    fa, bam = synth.write_dataset(str(tmp_path / "u"), cs)
    with hostlib.Bam(bam) as b:
        with pytest.raises(RuntimeError, match="no reads"):
            b.library_stats(20)
