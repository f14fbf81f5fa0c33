"""GPU path against the CPU oracle on one contig: every per-position array, duplicate flags, breakpoint clusters, candidate records,
gate events, the read-depth CNV state (mask, z list, window table, calls) and the record text -- bit for bit.

CHECKER ONLY (tests, __graft_entry__.smoke and bench.py's post-run `parity_check`); nothing on the product path imports this."""
from __future__ import annotations

import numpy as np


def compare_gpu_oracle(prm, contig, hez, mq, chr_name: str = None, cnv: bool = True, handle=None, device: int = 0) -> dict:
    """Runs `contig` (tools.synth.SynthContig) through the C ABI on the current device and through the oracle; raises AssertionError
    with the first difference; returns a few counts of what was compared."""
    from grom_b200 import gpu, hostlib
    from grom_b200.params import GA_NAMES
    from oracle import pyoracle as po

    name = (chr_name or contig.name).lower()
    if handle is None:
        gpu.init(device, hez, mq, prm)
    ch = handle if handle is not None else gpu.Chromosome(contig.batch.tid, contig.chars)
    try:
        if handle is not None:
            assert ch.rebind(contig.batch.tid, contig.chars), "the handle is too small for this contig"
        ch.push_reads(contig.batch)
        res = ch.finish()
        g = ch.cnv(params=prm) if cnv else None
        got = ch.fetch_all()
        cl = ch.fetch_clusters()
        state = ch.read_state(contig.batch.n_reads)
        z, mask = (ch.cnv_fetch("z"), ch.cnv_fetch("mask")) if cnv else (None, None)
    finally:
        if handle is None:
            ch.close()
    ref = po.run_chr(prm, contig.batch, contig.chars, hez, mq)
    assert np.array_equal(state, ref.read_state), "read_state (applied / -M duplicate) differs from the oracle"
    for k in range(len(GA_NAMES)):
        bad = np.nonzero(got[k] != ref.arrays[k])[0]
        assert bad.size == 0, f"array {GA_NAMES[k]}: {bad.size} mismatches, first at {bad[:5]}"
    w, rs, re, dist, mchr, ol = cl
    assert np.array_equal(w, ref.cl_w), "breakpoint cluster weights differ"
    live = ref.cl_w != 0
    assert np.array_equal(rs[live], ref.cl_rs[live]) and np.array_equal(re[live], ref.cl_re[live]), "cluster read_start / read_end differ"
    assert np.array_equal(dist[live], ref.cl_dist[live]), "cluster running-mean distances differ"
    assert np.array_equal(ol, ref.other_len), "other_len differs"
    assert (res.scan_first, res.scan_last) == (ref.scan_first, ref.scan_last), "scanned range differs"
    assert len(res.snv) == len(ref.snv) and np.array_equal(res.snv["pos"], ref.snv["pos"]), "SNV candidate set differs"
    for f in ("base", "ratio", "pr", "hez", "v"):
        assert np.array_equal(res.snv[f], ref.snv[f]), f"SNV candidate field {f} differs"
    assert res.snv_ave_rd == ref.snv_ave_rd or (np.isnan(res.snv_ave_rd) and np.isnan(ref.snv_ave_rd)), "SNV depth mean differs"
    assert np.array_equal(res.ins, ref.ins), "small-insertion candidates differ"
    assert np.array_equal(res.del_ev, ref.del_ev), "small-deletion scan events differ"
    assert res.sv_ev.tobytes() == ref.sv_ev.tobytes(), "structural-variant gate events differ"
    out = {"positions": int(len(contig.chars)), "reads": int(contig.batch.n_reads), "arrays": len(GA_NAMES), "snv": int(len(ref.snv)),
           "small_ins": int(len(ref.ins)), "small_del_events": int(len(ref.del_ev)), "sv_events": int(len(ref.sv_ev)),
           "dups": int((ref.read_state == 2).sum())}
    if cnv:
        o = po.cnv_run(prm, name, contig.chars, ref["gc"], ref["acgt"], ref["rd_mq"], ref["rd_rd"], ref["rd_low"], ploidy=prm.ploidy,
                       seed=prm.rand_seed, sample_cap=prm.sample_lists_len, min_win=prm.min_rd_window_len, max_win=prm.max_rd_window_len)
        assert np.array_equal(mask, o.mask), "CNV mask differs"
        assert np.array_equal(z, o.z), "CNV z list differs"
        assert np.array_equal(g.win_cnt, o.win_cnt) and np.array_equal(g.win_sd, o.win_sd), "CNV window table differs"
        dels, dups = g.calls[g.calls["kind"] == 0], g.calls[g.calls["kind"] == 1]
        for mine, want, what in ((dels, o.dels, "deletion"), (dups, o.dups, "duplication")):
            assert np.array_equal(mine["start"], want["start"]) and np.array_equal(mine["end"], want["end"]), f"read-depth {what} calls differ"
            for f, rf in (("z", "z"), ("cn", "cn"), ("cn_sd", "cs"), ("pvalue", "p")):
                assert np.array_equal(mine[f], want[rf]), f"read-depth {what} call field {f} differs"
        assert hostlib.vcf_cnv(prm, name, g.calls) == o.vcf, "read-depth CNV records differ"
        out.update(cnv_dels=int(len(o.dels)), cnv_dups=int(len(o.dups)))
        text = hostlib.vcf_contig(prm, name, contig.chars, res.snv, res.snv_ave_rd, res.ins, res.del_ev, res.sv_ev, g.calls)
        out["records"] = text.count("\n")
    return out
