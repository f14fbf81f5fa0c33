"""GPU parity under the flag combinations that round 1 only pinned CPU-side (tests/test_oracle_vs_reference.py, test_whole_vcf_live.py):
-q / -b, -v, read length 75, -p 4 -A 4 -g 1, and the bench's config-3 generator.  Everything goes through the C ABI and is compared
bit for bit with the oracle by tools/parity.py (arrays, -M flags, clusters, candidates, gate events, CNV mask / z / window table / calls)."""
import numpy as np
import pytest

from grom_b200 import gpu, hostlib
from grom_b200.params import Params
from tools import parity, synth, workloads

pytestmark = pytest.mark.gpu


def _spec(seed, read_len, length=400_000, depth=30, **kw):
    return synth.SynthSpec(contigs=[("chrA", length), ("chrB", 60_000)], depth=depth, seed=seed, read_len=read_len,
                           ins_mean=400.0 * read_len / 150, ins_sd=40.0 * read_len / 150, ins_floor=read_len + 20, dup_frac=0.05, clip_frac=0.04,
                           hardclip_frac=0.01, refskip_frac=0.003, disc_frac=0.03, sa_frac=0.8, munmap_frac=0.01, sv_sites_per_mb=10, sv_classes=6,
                           cnv_per_mb=2.0, cnv_min=15000, cnv_max=30000, **kw)


def _prm(read_len, **kw):
    im = int(400.0 * read_len / 150)
    kw = dict(kw)
    if "min_mapq" in kw:
        kw.setdefault("rd_min_mapq", kw["min_mapq"])                   # -q sets both (reference src/GROM.c:22102)
    if "pval_threshold" in kw:
        kw.setdefault("pval_threshold1", kw["pval_threshold"])         # -v sets both (src/GROM.c:22101)
    return Params.default(insert_mean=im, insert_min=read_len + 20, insert_max=im + int(120 * read_len / 150), lseq=read_len, **kw)


@pytest.mark.parametrize("seed,read_len,flags", [
    (51, 150, dict(rmdup=1, min_mapq=30, min_base_qual=26)),          # -M -q 30 -b 26
    (52, 100, dict(rmdup=0, min_mapq=4, min_base_qual=10)),           # -q 4 -b 10 (mq table of max(q,10))
    (53, 150, dict(rmdup=1, pval_threshold=0.01)),                    # -v 0.01
    (54, 75, dict(rmdup=1)),                                          # read length 75
    (55, 150, dict(rmdup=1, ploidy=4, windows_sampling_factor=4, gender=1)),   # -p 4 -A 4 -g 1
    (56, 250, dict(rmdup=0, min_mapq=30, min_base_qual=26, pval_threshold=0.01)),
])
def test_gpu_matches_oracle_under_flags(seed, read_len, flags):
    prm = _prm(read_len, **flags)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    for c in synth.simulate(_spec(seed, read_len)):
        out = parity.compare_gpu_oracle(prm, c, hez, mq)
        if c.name == "chrA":
            assert out["snv"] > 20 and out["sv_events"] > 0 and out["small_ins"] + out["small_del_events"] > 0
            if flags.get("rmdup"):
                assert out["dups"] > 0


def test_gpu_matches_oracle_config3_generator():
    """the bench's workload generator (tools/workloads.py) at 3 Mb: what bench.py's parity_check runs after its timed loops"""
    prm = workloads.params_config3()
    hez, mq = hostlib.tables(None, prm.min_mapq)
    c = synth.simulate(workloads.chr20_spec(mb=3, seed=77, cnv_per_mb=1.0))[0]
    b = c.batch
    assert (b.n_cigar > 1).sum() > 5000 and (b.sa_pos >= 0).sum() > 1000
    out = parity.compare_gpu_oracle(prm, c, hez, mq)
    assert out["small_ins"] > 0 and out["small_del_events"] > 0 and out["sv_events"] > 0 and out["cnv_dels"] + out["cnv_dups"] > 0


def test_result_buffers_grow_and_repeat(monkeypatch):
    """A contig whose evidence outgrows the first buffer sizes is run again with larger ones (no failure, same results): dense SV
    evidence on a short contig overflows the item buffer's first size (n / 4 + 65536 is generous, so shrink it through the env hook)."""
    prm = _prm(150, rmdup=1)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    c = synth.simulate(_spec(61, 150, length=300_000))[0]
    monkeypatch.setenv("GROMGPU_TEST_SMALL_BUFFERS", "1")
    out = parity.compare_gpu_oracle(prm, c, hez, mq)
    assert out["sv_events"] > 0


def test_out_of_order_reads_are_rejected():
    prm = _prm(150)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    c = synth.simulate(_spec(62, 150, length=100_000))[0]
    b = c.batch
    gpu.init(0, hez, mq, prm)
    pos = b.pos.copy()
    k = len(pos) // 2
    pos[k], pos[k + 40] = pos[k + 40], pos[k]
    assert pos[k] != pos[k + 40]
    b.pos = pos
    with gpu.Chromosome(0, c.chars) as ch:
        ch.push_reads(b)
        with pytest.raises(gpu.GromGpuError, match="out of coordinate order"):
            ch.run()


def test_lanes_on_second_device_if_present():
    """lane threads must run on the library's device, not on device 0.  Needs >= 2 GPUs; the 1-GPU box checks the same
    entry points from fresh threads (CUDA's current device is per thread)."""
    import ctypes, threading
    cuda = ctypes.CDLL("libcudart.so")
    n = ctypes.c_int(0)
    cuda.cudaGetDeviceCount(ctypes.byref(n))
    dev = 1 if n.value >= 2 else 0
    if dev == 1:
        pytest.skip("the library binds one device per process and earlier tests of this process initialised device 0")
    prm = _prm(150, rmdup=1)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    c = synth.simulate(_spec(63, 150, length=150_000))[0]
    gpu.init(dev, hez, mq, prm)
    outs, errs = [], []

    def lane():
        try:
            s = gpu.stream_create()
            with gpu.Chromosome(0, c.chars, stream=s) as ch:
                ch.push_reads(c.batch)
                r = ch.finish()
                g = ch.cnv(params=prm)
                outs.append((r.snv.tobytes(), r.sv_ev.tobytes(), g.calls.tobytes()))
            gpu.stream_destroy(s)
        except BaseException as e:
            errs.append(e)
    th = [threading.Thread(target=lane) for _ in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert len(outs) == 3 and outs[0] == outs[1] == outs[2]


def test_rebound_handle_matches_oracle():
    """One handle begun for the largest contig and rebound (gromgpu_chr_rebind: buffers kept, arrays zeroed) for shorter ones -- the
    per-lane flow of the genome drivers -- gives the oracle's arrays, candidates and read-depth calls on every contig; a longer contig
    is refused (return 1) and leaves the handle usable."""
    prm = _prm(150, rmdup=1)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    spec = _spec(61, 150, length=500_000)
    spec.contigs = [("chrA", 500_000), ("chrB", 320_000), ("chrC", 150_000), ("chrD", 40_000)]
    cs = synth.simulate(spec)
    gpu.init(0, hez, mq, prm)
    with gpu.Chromosome(cs[0].batch.tid, cs[0].chars) as ch:
        ch.push_reads(cs[0].batch); ch.finish(); ch.cnv(params=prm)          # leaves evidence, clusters and read-depth state of chrA behind
        for c in (cs[2], cs[1], cs[3], cs[0]):
            out = parity.compare_gpu_oracle(prm, c, hez, mq, handle=ch)
            assert out["positions"] == len(c.chars)
        big = np.full(600_000, ord("A"), dtype=np.uint8)
        assert ch.rebind(9, big) is False
        out = parity.compare_gpu_oracle(prm, cs[1], hez, mq, handle=ch)
        assert out["snv"] > 10
