"""Live cross-check of the read-depth CNV oracle (oracle/grom_oracle_cnv.c) against the white-box reference run here
(oracle/_ref/GROM_ref with dump hooks), in the modes no golden fixture pins: -p 4 -A 4 (BASELINE.json configs[4]) and a diploid
contig with planted (AT)n runs.  Every stage of src/GROM.c:16633-16990 and 18228-20355 is compared: pre-statistics, per-GC-bin
distributions, mask, z list, the 9,901-entry window sd table, the greedy calls with copy numbers, and the records.
Skipped where the binary is absent (the GPU box runs `-m gpu` only)."""
import os

import numpy as np
import pytest

from grom_b200.params import Params
from oracle import pyoracle as po
from tools import synth

pytestmark = pytest.mark.skipif(not po.have_reference("ref"), reason="oracle/_ref/GROM_ref not built")


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


@pytest.mark.parametrize("seed,length,depth,A,ploidy,at", [(7, 1_200_000, 20, 4, 4, 0), (9, 1_000_000, 30, 2, 2, 40)])
def test_live_reference_cnv_state_parity(tmp_path, seed, length, depth, A, ploidy, at):
    spec = synth.SynthSpec(contigs=[("chrA", length), ("chrZ", 50_000)], depth=depth, seed=seed, cnv_per_mb=4.0, disc_frac=0.005,
                           sv_sites_per_mb=1.0, low_mapq_frac=0.05, at_repeats=at, cnv_min=20_000, cnv_max=120_000)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), cs)
    dump = str(tmp_path / "dump")
    po.run_reference(bam, fa, str(tmp_path / "o.vcf"), args=["-A", A, "-p", ploidy, "-g", 1], dump_dir=dump, seed=5)
    m = po.read_mean_file(bam)
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"],
                         lseq=m["lseq"], windows_sampling_factor=A, ploidy=ploidy, gender=1)
    vcf = [l for l in open(str(tmp_path / "o.vcf")) if not l.startswith("#")]
    calls = 0
    for c in cs:
        name = c.name.lower()
        assert os.path.exists(os.path.join(dump, f"cnv_{name}.bin")), name
        dep = po.load_depth_dump(dump, name)
        gcd = po.load_gc_dump(dump, name)
        r = po.cnv_run(prm, name, c.chars, gcd[0], gcd[1], dep[0], dep[1], dep[2], ploidy=ploidy, seed=5)
        pre = po.load_cnvpre_dump(dump, name)
        d = po.load_cnv_dump(dump, name)
        for k in ("nblocks", "repeats", "chr_ave", "chr_sd", "rep_ave", "rep_sd", "rep_cnt", "biased", "blk_ave", "sample_blocks"):
            assert same(getattr(r, k), pre[k]), (name, "pre." + k)
        for k in ("z", "mask", "win_sd", "win_cnt", "ave", "sd", "del_thr", "dup_thr", "windows", "n_high", "n_low"):
            assert same(getattr(r, k), d[k]), (name, k)
        for k, mine in (("dels", r.dels), ("dups", r.dups)):
            for f in ("start", "end", "z", "cn", "cs"):
                assert same(mine[f], d[k][f]), (name, k, f)
            calls += len(mine)
        assert r.vcf.splitlines(keepends=True) == [l for l in vcf if l.startswith(name + "\t") and "\tSD:Z:CN:CS\t" in l], name
    assert calls >= 4                                   # the planted segments are found: the comparison is not vacuous
