"""GPU parity (run with -m gpu on the B200): everything goes through the C ABI (grom_b200/gpu.py is a thin
ctypes mirror of include/gromgpu.h).  Integer work => bit-exact comparisons."""
import os

import numpy as np
import pytest

from util import CHECKED, CLIPS, DEPTH, EVIDENCE, GCS, GOLDEN, PILEUP, assert_arrays_equal, assert_clusters_equal, golden_batches, golden_params, load_golden_fasta, tables, tables_7digit
from grom_b200 import gpu, hostlib
from grom_b200.params import GA, GA_NAMES, Params
from grom_b200.reads import CDEL, CHARD_CLIP, CINS, CMATCH, CREF_SKIP, CSOFT_CLIP, FPAIRED, FREVERSE, FMREVERSE
from oracle import pyoracle as po
from tools import synth

pytestmark = pytest.mark.gpu


def run_gpu(prm, batch_or_slices, fasta, hez, mq, tid=0):
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(tid, fasta) as ch:
        n = 0
        for b in (batch_or_slices if isinstance(batch_or_slices, list) else [batch_or_slices]):
            ch.push_reads(b); n += b.n_reads
        res = ch.finish()
        run_gpu.cnv = ch.cnv()
        run_gpu.clusters = ch.fetch_clusters()
        return res, ch.fetch_all(), ch.read_state(n), ch.stats()


def check_against_oracle(prm, batch, fasta, hez, mq, slices=None):
    res, got, state, st = run_gpu(prm, slices if slices else batch, fasta, hez, mq, tid=batch.tid)
    ref = po.run_chr(prm, batch, fasta, hez, mq)
    assert np.array_equal(state, ref.read_state)
    assert_arrays_equal(got, ref.arrays)
    assert_clusters_equal(run_gpu.clusters, ref)
    assert (res.scan_first, res.scan_last) == (ref.scan_first, ref.scan_last)
    assert np.array_equal(res.snv["pos"], ref.snv["pos"])
    for f in ("base", "ratio", "pr", "hez", "v"):
        assert np.array_equal(res.snv[f], ref.snv[f]), f
    assert len(res.ins) == len(ref.ins)
    for f in ("pos", "dist", "pr", "hez", "conc", "weight", "rd", "sc", "other_len", "seq"):
        assert np.array_equal(res.ins[f], ref.ins[f]), ("insertion candidate", f)
    assert np.array_equal(res.del_ev, ref.del_ev), "small-deletion scan events differ"
    assert len(res.sv_ev) == len(ref.sv_ev), ("structural-variant gate events", len(res.sv_ev), len(ref.sv_ev))
    assert res.sv_ev.tobytes() == ref.sv_ev.tobytes(), "structural-variant gate events differ"
    if len(ref.snv) or ref.scan_first >= 0:
        assert res.snv_ave_rd == ref.snv_ave_rd or (np.isnan(res.snv_ave_rd) and np.isnan(ref.snv_ave_rd))
    return res, got, state, st, ref


@pytest.mark.parametrize("tag,rmdup", [("default", 0), ("rmdup", 1)])
def test_gpu_reproduces_reference_golden(tag, rmdup):
    """CUDA path vs the fixtures dumped from the reference itself (not via the oracle)."""
    names, batches = golden_batches()
    fasta = load_golden_fasta()
    hez, mq = tables_7digit()
    g = np.load(os.path.join(GOLDEN, f"g1_{tag}.npz"))
    prm = golden_params(g, rmdup)
    vcf = str(g["vcf"]).splitlines(keepends=True)
    for tid, name in enumerate(names):
        n = name.lower()
        res, got, state, st = run_gpu(prm, batches[tid], fasta[name], hez, mq, tid=tid)
        pos, v = g[f"{n}_scan_pos"], g[f"{n}_scan_v"]
        assert (res.scan_first, res.scan_last) == (int(pos[0]), int(pos[-1]))
        for k in EVIDENCE:
            assert np.array_equal(got[k][pos], v[:, k]), (n, GA_NAMES[k])
        cw, crs, cre, cdist, cmchr, col = run_gpu.clusters
        d = g[f"{n}_scan_d"]
        for k in range(10):
            w_ref = v[:, 51 + 3 * k]; live = w_ref != 0
            assert np.array_equal(cw[k][pos], w_ref), (n, "cluster", k)
            assert np.array_equal(crs[k][pos][live], v[:, 52 + 3 * k][live]) and np.array_equal(cre[k][pos][live], v[:, 53 + 3 * k][live])
            assert np.array_equal(cdist[k][pos][live], d[:, k][live])
        assert np.array_equal(col[pos], v[:, 83])
        depth = g[f"{n}_depth"]
        for j, k in enumerate(DEPTH):
            assert np.array_equal(got[k], depth[j]), (n, GA_NAMES[k])
        gcd = g[f"{n}_gc"]
        M = prm.insert_mean
        lo, hi = M - 1, len(fasta[name]) - (2 * M - 1)
        assert np.array_equal(got[GA["gc"]][lo:hi], gcd[0][lo:hi]) and np.array_equal(got[GA["acgt"]][lo:hi], gcd[1][lo:hi])
        reads = g[f"{n}_reads"]
        proc = np.nonzero(state > 0)[0]
        assert np.array_equal(batches[tid].pos[proc], reads["pos"])
        assert np.array_equal((state[proc] == 1).astype(np.int32), reads["keep"])
        mine = po.format_snv_vcf(prm, n, fasta[name], res.snv, res.snv_ave_rd).splitlines(keepends=True)
        ref = [l for l in vcf if l.startswith(n + "\t") and l.split("\t")[2] == ""]
        assert mine == ref
        mine = po.normalise_records(hostlib.vcf_ins(prm, n, fasta[name], res.ins).splitlines(keepends=True))
        ref = po.normalise_records([l for l in vcf if l.startswith(n + "\t") and "\tSPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP\t" in l])
        assert len(ref) > 0 and mine == ref
        # product host writer on the GPU results == the reference's records, all three classes
        assert hostlib.vcf_snv(prm, n, fasta[name], res.snv, res.snv_ave_rd).splitlines(keepends=True) == [l for l in vcf if l.startswith(n + "\t") and l.split("\t")[2] == ""]
        mine = hostlib.vcf_smalldel(prm, n, fasta[name], res.del_ev).splitlines(keepends=True)
        assert mine == [l for l in vcf if l.startswith(n + "\t") and "\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT:SSC:ESC:HP\t" in l]
        # and the complete record text of the contig (all classes incl. SV merge and read-depth CNV), in the reference's order
        full = hostlib.vcf_contig(prm, n, fasta[name], res.snv, res.snv_ave_rd, res.ins, res.del_ev, res.sv_ev, run_gpu.cnv.calls)
        assert po.normalise_records(full.splitlines(keepends=True)) == po.normalise_records([l for l in vcf if l.startswith(n + "\t")])


@pytest.mark.parametrize("seed,rmdup,read_len,depth", [(1, 0, 150, 30), (2, 1, 150, 30), (3, 1, 100, 60), (4, 0, 250, 10)])
def test_gpu_matches_oracle_synthetic(seed, rmdup, read_len, depth):
    spec = synth.SynthSpec(contigs=[("chrA", 300_000), ("chrB", 100_000)], depth=depth, seed=seed, read_len=read_len,
                           ins_mean=2.8 * read_len, ins_sd=35, ins_floor=read_len + 15, dup_frac=0.06, clip_frac=0.04,
                           hardclip_frac=0.01, refskip_frac=0.003, long_name_frac=0.01, disc_frac=0.03, sa_frac=0.8, munmap_frac=0.01,
                           sv_sites_per_mb=10)
    hez, mq = tables()
    prm = Params.default(insert_mean=int(2.8 * read_len), insert_min=read_len + 15, insert_max=int(2.8 * read_len) + 120,
                         lseq=read_len, rmdup=rmdup)
    for c in synth.simulate(spec):
        res, got, state, st, ref = check_against_oracle(prm, c.batch, c.chars, hez, mq)
        assert len(res.snv) > 20
        if rmdup:
            assert st.n_dups == int((ref.read_state == 2).sum()) > 0


def test_multi_push_equals_single_push():
    spec = synth.SynthSpec(contigs=[("chrA", 150_000)], depth=20, seed=8, dup_frac=0.05)
    c = synth.simulate(spec)[0]
    hez, mq = tables()
    prm = Params.default(insert_min=170, insert_max=520, rmdup=1)
    n = c.batch.n_reads
    cuts = [0, n // 5, n // 5 + 1, n // 2, n]
    slices = [synth.slice_batch(c.batch, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    check_against_oracle(prm, c.batch, c.chars, hez, mq, slices=slices)


def test_transport_compact_forms_equal_canonical_upload():
    """include/grom_reads.h GROM_LAYOUT_*: offsets derived on the device, 2-bit / 4-bit dictionary qualities, 2-bit bases with an exception list, sparse SA fields --
    pushed whole and in slices -- give the oracle's arrays, flags, candidates and events like the canonical arrays do."""
    from grom_b200.reads import LAYOUT_CANONICAL_OFFSETS, LAYOUT_QUAL2, LAYOUT_QUAL4, LAYOUT_SEQ2, LAYOUT_SPARSE_SA
    from test_reads_compact import more_qualities, plant_non_acgt
    spec = synth.SynthSpec(contigs=[("chrA", 200_000)], depth=25, seed=17, dup_frac=0.05, clip_frac=0.04, hardclip_frac=0.01, refskip_frac=0.003,
                           disc_frac=0.03, sa_frac=0.8, munmap_frac=0.01, sv_sites_per_mb=10)
    c = synth.simulate(spec)[0]
    hez, mq = tables()
    prm = Params.default(insert_min=170, insert_max=520, rmdup=1)
    b = plant_non_acgt(c.batch.repack_canonical(), n_runs=2000).compact()
    assert b.layout_flags == LAYOUT_CANONICAL_OFFSETS | LAYOUT_QUAL2 | LAYOUT_SPARSE_SA | LAYOUT_SEQ2
    assert 0 < len(b.sa_index) < b.n_reads and len(b.seq_exc_slot) > 5000
    check_against_oracle(prm, b, c.chars, hez, mq)
    n = b.n_reads
    cuts = [0, n // 7, n // 7 + 3, n // 2, n]
    slices = [synth.slice_batch(c.batch, a, e).compact() for a, e in zip(cuts[:-1], cuts[1:])]
    slices[2].layout_flags &= ~LAYOUT_QUAL2                       # mixed forms across pushes
    slices[3].layout_flags &= ~LAYOUT_SEQ2
    slices[1].layout_flags = 0
    check_against_oracle(prm, c.batch, c.chars, hez, mq, slices=slices)
    # nine distinct qualities: the 4-bit dictionary
    c.batch.layout_flags = 0; c.batch.qual2 = None
    b4 = more_qualities(c.batch).compact()
    assert (b4.layout_flags & LAYOUT_QUAL4) and not (b4.layout_flags & LAYOUT_QUAL2)
    check_against_oracle(prm, b4, c.chars, hez, mq)
    # more than 16 distinct qualities: the dictionary form is not offered, the rest still is
    c.batch.layout_flags = 0; c.batch.qual4 = None
    c.batch.qual[: 40] = np.arange(40, dtype=np.uint8) + 2
    b2 = c.batch.compact()
    assert not (b2.layout_flags & (LAYOUT_QUAL4 | LAYOUT_QUAL2)) and (b2.layout_flags & LAYOUT_SPARSE_SA)
    check_against_oracle(prm, b2, c.chars, hez, mq)


def _ref(n, seed=0):
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()


def test_empty_and_all_skipped_inputs():
    hez, mq = tables()
    prm = Params.default()
    fa = _ref(20_000)
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(0, fa) as ch:          # no reads at all
        res = ch.finish()
        assert (res.scan_first, res.scan_last, len(res.snv)) == (-1, -1, 0)
        assert not ch.fetch_all()[EVIDENCE + DEPTH].any()
    # every read before W/4+1: nothing is applied, nothing is scanned (reference src/GROM.c:6406)
    recs = [dict(pos=p, cigar=[(CMATCH, 50)], seq="ACGTA" * 10) for p in (10, 500, 3000)]
    b = synth.batch_from_records(0, recs)
    res, got, state, st, ref = check_against_oracle(prm, b, fa, hez, mq)
    assert res.scan_first == -1 and not state.any() and not got[EVIDENCE + DEPTH].any()


def test_edge_cigars_and_ragged_reads():
    """Ragged lengths, leading/trailing H and S, I/D/N, '=' and 'X', a >1000-op CIGAR (pileup cap src/GROM.c:6741,
    depth walk uncapped src/GROM.c:6615), N base calls, IUPAC reference letters, a read hanging over the contig end."""
    hez, mq = tables()
    prm = Params.default()
    P = 30_000
    fa = _ref(P, 3)
    fa[9000:9010] = np.frombuffer(b"NNNNRYKMnn", dtype=np.uint8)
    fa[12000:12100] |= 0x20
    s = lambda n, k=0: "".join("ACGT"[(i * 7 + k) % 4] for i in range(n))  # noqa: E731
    many = [(CMATCH, 1), (CINS, 1)] * 600 + [(CMATCH, 5)]
    recs = [
        dict(pos=8990, cigar=[(CMATCH, 40)], seq="N" * 5 + s(35), qual=list(range(10, 50))),
        dict(pos=9000, cigar=[(CHARD_CLIP, 7), (CSOFT_CLIP, 3), (CMATCH, 20), (CINS, 4), (CMATCH, 10), (CDEL, 6), (CMATCH, 13), (CSOFT_CLIP, 5), (CHARD_CLIP, 2)],
             seq=s(55, 1), flag=FREVERSE),
        dict(pos=9005, cigar=[(7, 12), (8, 1), (CREF_SKIP, 300), (CMATCH, 17)], seq=s(30, 2), mapq=5),
        dict(pos=9100, cigar=many, seq=s(1205, 3), qual=25),
        dict(pos=12010, cigar=[(CMATCH, 1)], seq="A"),
        dict(pos=12010, cigar=[(CMATCH, 250)], seq=s(250), qual=19),
        dict(pos=P - 30, cigar=[(CMATCH, 50)], seq=s(50)),
        dict(pos=P - 30, cigar=[(CMATCH, 20), (CDEL, 3), (CMATCH, 30)], seq=s(50, 1)),
        dict(pos=P - 1, cigar=[(CMATCH, 1)], seq="G"),
    ]
    b = synth.batch_from_records(0, recs)
    res, got, state, st, ref = check_against_oracle(prm, b, fa, hez, mq)
    assert got[GA["rd_rd"]].sum() > 0 and got[GA["snvlow_a"]].sum() > 0


def test_deep_pile_same_position_duplicates_and_name_slots():
    """4,000 reads starting at one position: the -M run scan, >3 distinct mismatching names, mates that overlap
    with identical names (src/GROM.c:6805-6824), long names that are never stored."""
    hez, mq = tables()
    prm = Params.default(rmdup=1)
    fa = np.full(20_000, ord("A"), dtype=np.uint8)
    rng = np.random.default_rng(5)
    recs = []
    for i in range(4000):
        paired = i % 3 != 0
        recs.append(dict(pos=8000, cigar=[(CMATCH, 60)], seq="".join(rng.choice(list("ACGT"), 60, p=[0.7, 0.1, 0.1, 0.1])),
                         qual=int(rng.integers(10, 41)), mapq=int(rng.choice([60, 60, 10])),
                         flag=(FPAIRED | (FREVERSE if i % 2 else FMREVERSE)) if paired else 0,
                         mpos=8000 + int(rng.integers(0, 4)) * 100, tlen=int(rng.integers(0, 3)) * 10,
                         name=("q%d" % (i % 700)) if i % 11 else ("L" * 60 + str(i % 5))))
    b = synth.batch_from_records(0, recs)
    res, got, state, st, ref = check_against_oracle(prm, b, fa, hez, mq)
    assert (state == 2).sum() > 100


def test_unsorted_push_is_rejected():
    hez, mq = tables()
    gpu.init(0, hez, mq, Params.default())
    fa = _ref(20_000)
    b1 = synth.batch_from_records(0, [dict(pos=9000, cigar=[(CMATCH, 10)], seq="ACGTACGTAC")])
    b2 = synth.batch_from_records(0, [dict(pos=8000, cigar=[(CMATCH, 10)], seq="ACGTACGTAC")])
    with gpu.Chromosome(0, fa) as ch:
        ch.push_reads(b1)
        with pytest.raises(gpu.GromGpuError, match="coordinate order"):
            ch.push_reads(b2)


def test_size_independent_properties_large():
    """4 Mb / 30x (0.8 M reads): checksums that hold at any size, plus idempotence of a re-run."""
    spec = synth.SynthSpec(contigs=[("chrL", 4_000_000)], depth=30, seed=77, simple=True, ins_floor=310)
    c = synth.simulate(spec)[0]
    hez, mq = tables()
    prm = Params.default(insert_min=310, insert_max=520)
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(0, c.chars) as ch:
        ch.push_reads(c.batch)
        r1 = ch.finish(); a1 = ch.fetch_all(); st = ch.stats()
        r2 = ch.finish(); a2 = ch.fetch_all()
    assert np.array_equal(a1[CHECKED], a2[CHECKED]) and np.array_equal(r1.snv, r2.snv)
    applied = c.batch.pos >= prm.first_pos
    inside = applied & (c.batch.pos + 150 < len(c.chars))
    # every aligned base of an applied read lands in exactly one depth counter (src/GROM.c:6621-6658)
    assert int(a1[GA["rd_rd"]].sum(dtype=np.int64) + a1[GA["rd_low"]].sum(dtype=np.int64)) == int(inside.sum()) * 150
    assert int(a1[GA["rd_mq"]].sum(dtype=np.int64)) == int(c.batch.mapq[inside].astype(np.int64).sum()) * 150
    # ins_floor >= 2*read_len => no mate overlap => no name skip: every A/C/G/T base is counted once in snv or snv_lowmq
    tot = sum(int(a1[k].sum(dtype=np.int64)) for k in range(8))
    nb = 0
    for i in np.nonzero(applied)[0][:: max(1, applied.sum() // 2000)]:
        pass
    assert int(a1[GA["rc_all"]].sum(dtype=np.int64)) == tot
    assert np.array_equal(a1[GA["bq_rc"]], a1[0] + a1[1] + a1[2] + a1[3])
    assert st.aligned_bases == int(applied.sum()) * 150
    # physical depth: sum of rd == read spans + insert gaps of concordant forward reads (src/GROM.c:7176-7181, 8345-8365)
    assert int(a1[GA["rd"]].sum(dtype=np.int64)) >= int(np.minimum(c.batch.pos[applied] + 150, len(c.chars)).sum() - c.batch.pos[applied].sum())
    assert int(a1[GA["conc"]].sum(dtype=np.int64)) > 0 and np.all(a1[GA["conc"]] <= a1[GA["rd"]])
    # planted homozygous SNVs far from the ends are all called
    hom = c.truth["snv_pos"][~c.truth["snv_het"]]
    hom = hom[(hom > prm.first_pos + 200) & (hom < r1.scan_last - 200)]
    called = np.isin(hom, r1.snv["pos"])
    assert called.mean() > 0.97


def test_sv_gate_events_all_classes():
    """Planted clusters of every class (deletion, tandem duplication, both inversion orientations, translocation, insertion): the gate
    events of the CUDA path equal the oracle's, and the host list builder turns both into the same candidate lists."""
    spec = synth.SynthSpec(contigs=[("chrA", 400_000), ("chrB", 150_000)], depth=30, seed=14, dup_frac=0.05, sa_frac=0.5, disc_frac=0.03,
                           sv_sites_per_mb=10.0, munmap_frac=0.01, sv_classes=25)
    cs = synth.simulate(spec)
    prm = Params.default(insert_mean=400, insert_min=150, insert_max=530, lseq=150)
    hez, mq = tables_7digit()
    seen = set()
    for c in cs:
        res = check_against_oracle(prm, c.batch, c.chars, hez, mq)[0]
        seen |= set(res.sv_ev["cls"].tolist())
        a = hostlib.sv_lists(prm, res.sv_ev)
        assert sum(len(v) for v in a.values()) > 0
        ref = po.run_chr(prm, c.batch, c.chars, hez, mq)
        ocnv = po.cnv_run(prm, c.name.lower(), c.chars, ref["gc"], ref["acgt"], ref["rd_mq"], ref["rd_rd"], ref["rd_low"])
        full_gpu = hostlib.vcf_contig(prm, c.name.lower(), c.chars, res.snv, res.snv_ave_rd, res.ins, res.del_ev, res.sv_ev, run_gpu.cnv.calls)
        full_cpu = hostlib.vcf_contig(prm, c.name.lower(), c.chars, ref.snv, ref.snv_ave_rd, ref.ins, ref.del_ev, ref.sv_ev, run_gpu.cnv.calls)
        assert full_gpu == full_cpu and "<DUP>" in full_gpu and "<INV>" in full_gpu and "<DEL>" in full_gpu
        assert len(run_gpu.cnv.calls) == len(ocnv.dels) + len(ocnv.dups)
        assert (a["dup"]["end"]["pos"] >= 0).any() and (a["del"]["end"]["pos"] >= 0).any()     # pairs were completed
    assert seen == set(range(12)), seen
