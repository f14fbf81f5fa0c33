"""Host batcher probe: writes (once) a BAM of the config-3 generator and times gromhost_bam_read_target on it.
GROMHOST_TRACE=1 makes the library print its phase times.  Nothing here is on the product path."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grom_b200 import hostlib  # noqa: E402
from tools import synth, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=4.0)
ap.add_argument("--depth", type=float, default=30.0)
ap.add_argument("--dir", default="/tmp/grom_decode_probe")
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true", help="compare the decoded batch with the generator's arrays and with the all-zlib decode")
a = ap.parse_args()
os.makedirs(a.dir, exist_ok=True)
stem = os.path.join(a.dir, f"p{a.mb:g}_{a.depth:g}")
cs = None
if not os.path.exists(stem + ".bam") or a.check:
    t0 = time.perf_counter()
    cs = synth.simulate(workloads.chr20_spec(mb=a.mb, depth=a.depth, seed=2020, name="chr20p", names=True, cnv_per_mb=0.5))
    synth.write_dataset(stem, cs)
    print(f"generated {stem}.bam in {time.perf_counter() - t0:.1f} s, {os.cpu_count()} cores", flush=True)
best = None
for _ in range(a.reps):
    t0 = time.perf_counter()
    with hostlib.Bam(stem + ".bam") as bf:
        bt = bf.read_target_owned(0, threads=a.threads)       # the product call: the batch stays in the batcher's memory
    dt = time.perf_counter() - t0
    bt.free()
    best = dt if best is None else min(best, dt)
    print(f"{dt * 1e3:8.1f} ms  {bt.n_reads} reads", flush=True)
print(f"best {best * 1e3:.1f} ms = {bt.n_reads / best / 1e6:.2f} M reads/s, BAM {os.path.getsize(stem + '.bam') / 1e6:.1f} MB")
if a.check:
    import numpy as np
    with hostlib.Bam(stem + ".bam") as bf:
        mine = bf.read_target(0, threads=a.threads)
        os.environ["GROMHOST_INFLATE"] = "zlib"
        theirs = bf.read_target(0, threads=a.threads)
    o = cs[0].batch
    bad = []
    for k in ("pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len", "qname_hash", "cigar", "sa_pos", "sa_strand", "sa_mapq",
              "sa_same_chr", "sa_start_adj", "sa_end_adj", "sa_end_adj_indel"):
        if not np.array_equal(getattr(mine, k), getattr(o, k)):
            bad.append(k)
    for i in range(0, mine.n_reads, 997):
        if not (np.array_equal(mine.bases(i), o.bases(i)) and np.array_equal(mine.quals(i), o.quals(i))):
            bad.append(f"bases/quals of read {i}")
            break
    for k in ("seq4", "qual", "seq2", "qual2", "seq_exc_slot", "seq_exc_code", "sa_index", "cigar_off", "base_off"):
        if not np.array_equal(getattr(mine, k), getattr(theirs, k)):
            bad.append("zlib decode differs: " + k)
    print("check:", "ok" if not bad else "DIFFERENT " + ", ".join(bad), f"({mine.n_reads} reads, layout flags {mine.layout_flags})")
    sys.exit(1 if bad else 0)
