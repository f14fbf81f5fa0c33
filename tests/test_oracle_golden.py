"""The CPU oracle against fixtures produced by the reference itself (tests/golden/make_golden.py): per-position
count arrays at every scanned position, per-read -M keep flags, CNV depth arrays, scan range, look-ahead read
length and the SNV VCF records."""
import os

import numpy as np
import pytest

from util import GOLDEN, CHECKED, PILEUP, CLIPS, golden_batches, golden_params, load_golden_fasta, tables_7digit
from grom_b200.params import GA, GA_NAMES
from oracle import pyoracle as po


@pytest.fixture(scope="module")
def data():
    names, batches = golden_batches()
    return names, batches, load_golden_fasta(), tables_7digit()


@pytest.mark.parametrize("tag,rmdup", [("default", 0), ("rmdup", 1)])
def test_oracle_reproduces_reference_dumps(data, tag, rmdup):
    names, batches, fasta, (hez, mq) = data
    g = np.load(os.path.join(GOLDEN, f"g1_{tag}.npz"))
    prm = golden_params(g, rmdup)
    vcf = str(g["vcf"]).splitlines(keepends=True)
    for tid, name in enumerate(names):
        n = name.lower()
        r = po.run_chr(prm, batches[tid], fasta[name], hez, mq)
        pos = g[f"{n}_scan_pos"]
        v = g[f"{n}_scan_v"]
        assert (r.scan_first, r.scan_last) == (int(pos[0]), int(pos[-1]))
        assert np.array_equal(pos, np.arange(pos[0], pos[-1] + 1))
        for k in range(51):                                   # every per-position int the reference keeps (oracle/hooks.h)
            bad = np.nonzero(r.arrays[k][pos] != v[:, k])[0]
            assert bad.size == 0, (n, GA_NAMES[k], pos[bad[:5]])
        d = g[f"{n}_scan_d"]
        for k in range(10):                                   # the ten breakpoint clusters: weight, first/last read, running mean
            w_ref = v[:, 51 + 3 * k]
            assert np.array_equal(r.cl_w[k][pos], w_ref), (n, "cluster weight", k)
            live = w_ref != 0
            assert np.array_equal(r.cl_rs[k][pos][live], v[:, 52 + 3 * k][live]), (n, "read_start", k)
            assert np.array_equal(r.cl_re[k][pos][live], v[:, 53 + 3 * k][live]), (n, "read_end", k)
            assert np.array_equal(r.cl_dist[k][pos][live], d[:, k][live]), (n, "dist", k)      # doubles, bit-exact
        for k in range(2):
            live = v[:, 51 + 3 * (8 + k)] != 0
            assert np.array_equal(r.cl_mchr[k][pos][live], v[:, 81 + k][live])
        assert np.array_equal(r.other_len[pos], v[:, 83])
        assert np.array_equal(r.lookahead_lseq[pos], v[:, 84])
        depth = g[f"{n}_depth"]
        for j, k in enumerate(("rd_mq", "rd_rd", "rd_low")):
            assert np.array_equal(r[k], depth[j]), (n, k)
        gcd = g[f"{n}_gc"]
        M = prm.insert_mean
        lo, hi = M - 1, len(fasta[name]) - (2 * M - 1)      # the reference writes only this range (src/GROM.c:1684)
        assert np.array_equal(r["gc"][lo:hi], gcd[0][lo:hi]) and np.array_equal(r["acgt"][lo:hi], gcd[1][lo:hi])
        reads = g[f"{n}_reads"]
        proc = np.nonzero(r.read_state > 0)[0]
        assert len(proc) == len(reads)
        assert np.array_equal(batches[tid].pos[proc], reads["pos"])
        assert np.array_equal((r.read_state[proc] == 1).astype(np.int32), reads["keep"])
        if rmdup:
            assert (reads["keep"] == 0).sum() > 0
        mine = po.format_snv_vcf(prm, n, fasta[name], r.snv, r.snv_ave_rd).splitlines(keepends=True)
        ref = [l for l in vcf if l.startswith(n + "\t") and l.split("\t")[2] == ""]
        assert len(ref) > 5 and mine == ref
        mine = po.normalise_records(po.format_ins_vcf(prm, n, fasta[name], r.ins).splitlines(keepends=True))
        ref = po.normalise_records([l for l in vcf if l.startswith(n + "\t") and "\tSPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP\t" in l])
        assert len(ref) > 0 and mine == ref          # small-insertion records (ECO/EOT are uninitialised in the reference)


def test_reference_ignores_reads_before_quarter_window(data):
    names, batches, fasta, (hez, mq) = data
    g = np.load(os.path.join(GOLDEN, "g1_default.npz"))
    prm = golden_params(g, 0)
    r = po.run_chr(prm, batches[0], fasta[names[0]], hez, mq)
    assert r.scan_first == prm.window_len // 4 + 1
    assert np.all(r.read_state[batches[0].pos < r.scan_first] == 0)
