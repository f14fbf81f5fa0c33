"""Quick device-side timing of one synthetic contig (development aid, not the bench)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grom_b200 import gpu, hostlib
from grom_b200.params import Params
from tools import synth

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=8.0)
ap.add_argument("--depth", type=float, default=30.0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--simple", type=int, default=1)
ap.add_argument("--rmdup", type=int, default=1)
ap.add_argument("--cnv-per-mb", type=float, default=0.0)
ap.add_argument("--skip-e2e", type=int, default=0)
ap.add_argument("--skip-cnv", type=int, default=0)
ap.add_argument("--workload", default="config3", choices=["config3", "simple"], help="config3 = tools/workloads.chr20_spec (what bench.py runs)")
ap.add_argument("--ploidy", type=int, default=2)
ap.add_argument("--A", type=int, default=2)
ap.add_argument("--seed", type=int, default=20)
ap.add_argument("--dummy", type=int, default=1_000_000)
a = ap.parse_args()
t = time.time()
if a.workload == "config3":
    from tools import workloads
    spec = workloads.chr20_spec(mb=a.mb, depth=a.depth, seed=a.seed, name="chrP", cnv_per_mb=a.cnv_per_mb if a.cnv_per_mb > 0 else 0.25, dummy_len=a.dummy)
else:
    spec = synth.SynthSpec(contigs=[("chrP", int(a.mb * 1e6))], depth=a.depth, seed=20, simple=bool(a.simple), dup_frac=0.05, simple_disc_frac=0.01, names=False, cnv_per_mb=a.cnv_per_mb)
c = synth.simulate(spec)[0]
print("generated", c.batch.n_reads, "reads in", round(time.time() - t, 1), "s", flush=True)
prm = Params.default(insert_mean=400, insert_min=170, insert_max=520, lseq=150, rmdup=a.rmdup, ploidy=a.ploidy, windows_sampling_factor=a.A)
hez, mq = hostlib.tables(None, prm.min_mapq)
gpu.init(0, hez, mq, prm)
with gpu.Chromosome(0, c.chars) as ch:
    t = time.time(); ch.push_reads(c.batch); ch.run(); print("first run incl H2D", round(time.time() - t, 3), "s")
    for r in range(a.reps):
        ch.run()
        s = ch.stats().as_dict()
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items()}))
    P = len(c.chars)
    s = ch.stats()
    alg = s.bytes_reads + 104 * P   # 26 arrays written by the pileup kernel
    print(f"pileup: {alg/1e9:.3f} GB algorithmic / {s.ms_pileup:.3f} ms = {alg/s.ms_pileup/1e6:.1f} GB/s; "
          f"aligned bases/s (device total) = {s.aligned_bases/s.ms_total/1e-3/1e9:.2f} G")
    for r in range(0 if a.skip_cnv else 2):
        t = time.time(); g = ch.cnv(); dt = time.time() - t
        print(json.dumps(dict(cnv_ms_total=round(g.ms_total, 2), cnv_ms_device=round(g.ms_device, 2), cnv_ms_host=round(g.ms_host, 2), wall_ms=round(dt * 1e3, 2),
                              calls=len(g.calls), samples=g.n_samples, frames=g.n_frames, repeats=g.n_repeats, biased=g.biased_repeat)))
    # end-to-end breakdown with pageable-vs-pinned note: the bench pins its batch; here the arrays are numpy (pageable)
    if a.skip_e2e: sys.exit(0)
    import ctypes
    cuda = ctypes.CDLL("libcudart.so")
    def sync(): cuda.cudaDeviceSynchronize()
    for r in range(2):
        sync(); t0 = time.time(); ch.reset(c.chars); sync(); t1 = time.time(); ch.push_reads(c.batch); sync(); t2 = time.time()
        ch.run(); sync(); t3 = time.time(); res = ch.result(); t4 = time.time(); g = ch.cnv(); t5 = time.time()
        print(json.dumps(dict(e2e_ms=dict(reset=round((t1 - t0) * 1e3, 2), push=round((t2 - t1) * 1e3, 2), run=round((t3 - t2) * 1e3, 2),
                                          result=round((t4 - t3) * 1e3, 2), cnv=round((t5 - t4) * 1e3, 2)))))
