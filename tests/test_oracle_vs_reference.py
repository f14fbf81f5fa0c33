"""Live cross-check (build container only): run the white-box reference (oracle/_ref/GROM_ref) on freshly
generated data and compare the oracle with its dumps.  Skipped where the binary is absent."""
import os

import numpy as np
import pytest

from util import CLIPS, PILEUP
from grom_b200 import hostlib
from grom_b200.params import GA_NAMES, Params
from oracle import pyoracle as po
from tools import synth

pytestmark = pytest.mark.skipif(not po.have_reference("ref"), reason="oracle/_ref/GROM_ref not built")


@pytest.mark.parametrize("seed,rmdup,read_len,q,bq", [(31, 0, 150, 20, 20), (32, 1, 100, 20, 20), (33, 1, 75, 20, 20), (34, 0, 250, 20, 20),
                                                       (35, 1, 150, 30, 26), (36, 0, 100, 4, 10)])
def test_live_reference_dump_parity(tmp_path, seed, rmdup, read_len, q, bq):
    spec = synth.SynthSpec(contigs=[("chrQ", 120_000), ("chrR", 50_000), ("chrZ", 10_000)], depth=25, seed=seed,
                           read_len=read_len, ins_mean=3.0 * read_len, ins_sd=30, ins_floor=read_len + 20,
                           dup_frac=0.05, clip_frac=0.04, disc_frac=0.03, sa_frac=0.8, munmap_frac=0.01)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), cs)
    dump = str(tmp_path / "dump")
    po.run_reference(bam, fa, str(tmp_path / "o.vcf"), args=(["-M"] if rmdup else []) + (["-q", q, "-b", bq] if (q, bq) != (20, 20) else []), dump_dir=dump)
    m = po.read_mean_file(bam)
    prm = Params.default(insert_mean=max(m["insert_mean"], m["lseq"]), insert_min=m["insert_min"], insert_max=m["insert_max"],
                         lseq=m["lseq"], rmdup=rmdup, min_mapq=q, rd_min_mapq=q, min_base_qual=bq)   # -q sets both, src/GROM.c:22102
    hez, mq = po.reference_tables(q)
    vcf = [l for l in open(str(tmp_path / "o.vcf")) if not l.startswith("#")]
    with hostlib.Bam(bam) as b:
        for tid, c in enumerate(cs):
            n = c.name.lower()
            batch = b.read_target(tid)
            r = po.run_chr(prm, batch, c.chars, hez, mq)
            sd = po.load_scan_dump(dump, n)
            pos = sd["pos"]
            assert (r.scan_first, r.scan_last) == (int(pos[0]), int(pos[-1]))
            for k in range(51):
                assert np.array_equal(r.arrays[k][pos], sd["v"][:, k]), (n, GA_NAMES[k])
            for k in range(10):
                w_ref = sd["v"][:, 51 + 3 * k]
                live = w_ref != 0
                assert np.array_equal(r.cl_w[k][pos], w_ref)
                assert np.array_equal(r.cl_rs[k][pos][live], sd["v"][:, 52 + 3 * k][live])
                assert np.array_equal(r.cl_re[k][pos][live], sd["v"][:, 53 + 3 * k][live])
                assert np.array_equal(r.cl_dist[k][pos][live], sd["d"][:, k][live])
            assert np.array_equal(r.other_len[pos], sd["v"][:, 83])
            assert np.array_equal(r.lookahead_lseq[pos], sd["v"][:, 84])
            dd = po.load_depth_dump(dump, n)
            for j, k in enumerate(("rd_mq", "rd_rd", "rd_low")):
                assert np.array_equal(r[k], dd[j])
            rd = po.load_reads_dump(dump, n)
            proc = np.nonzero(r.read_state > 0)[0]
            assert np.array_equal(batch.pos[proc], rd["pos"])
            assert np.array_equal((r.read_state[proc] == 1).astype(np.int32), rd["keep"])
            mine = po.format_snv_vcf(prm, n, c.chars, r.snv, r.snv_ave_rd).splitlines(keepends=True)
            ref = [l for l in vcf if l.startswith(n + "\t") and l.split("\t")[2] == ""]
            assert mine == ref
