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
a = ap.parse_args()
os.makedirs(a.dir, exist_ok=True)
stem = os.path.join(a.dir, f"p{a.mb:g}_{a.depth:g}")
if not os.path.exists(stem + ".bam"):
    cs = synth.simulate(workloads.chr20_spec(mb=a.mb, depth=a.depth, seed=2020, name="chr20p", names=True, cnv_per_mb=0.5))
    synth.write_dataset(stem, cs)
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
