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
    assert {k: st[k] for k in KEYS} == {k: ref[k] for k in KEYS}
