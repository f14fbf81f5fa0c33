#!/usr/bin/env python
"""bench.py -- aligned bases/sec through the GROM hot path on B200 (BASELINE.json metric).

One "step" = one pass of the whole hot path (duplicate flags, read prep + clip scatter, pileup + CNV depth,
range-add scan, SNV gate + compaction) over one chr20-sized synthetic contig (64 Mb, 30x, 2x150 bp,
BASELINE.json configs[2]) per GPU.  Ranks are independent (chromosomes partition across GPUs exactly like the
reference's -P processes, no data-path collective): weak scaling, value = bases of all ranks / max-over-ranks time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mb 64] [--depth 30]
  python bench.py --impl reference ...      times the reference's own CPU implementation (oracle/_ref/GROM_ref,
                                            built from the reference's translation unit) with -P on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned_bases_per_sec_hot_path"
UNIT = "bases/s"
N_ARRAYS_PILEUP = 26          # int32 arrays the pileup kernel writes per position (23 pileup + 3 CNV depth)
N_ARRAYS_PILEUP_READ = 2      # rd and indel_sc_rd, read by the SNV gate fused into the pileup epilogue


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mb", type=float, default=64.0, help="contig length per GPU in Mb (default: chr20-sized)")
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--cpu-sample-mb", type=float, default=1.5, help="contig length of each CPU-baseline sample contig")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="N > 1: do not give every rank its own cores near its GPU (bind_rank_to_cores)")
    ap.add_argument("--lanes", type=int, default=3, help="contigs in flight in the end-to-end measurement")
    ap.add_argument("--canonical-upload", action="store_true", help="upload the canonical arrays only (no transport-compact forms)")
    ap.add_argument("--parity-mb", type=float, default=4.0, help="length of the contig (same generator) checked against the oracle after the timed loops; 0 = skip")
    ap.add_argument("--simple", action="store_true", help="round-1 workload: single-M reads only (kernel best case; not config 3)")
    ap.add_argument("--cnv-per-mb", type=float, default=0.25)
    ap.add_argument("--workload", default="chr20", choices=["chr20", "tetra", "wgs"],
                    help="chr20 = config 3 (default, the metric's configuration); tetra = config 5 (100x, -p 4 -A 4, one 64 Mb contig per GPU); "
                         "wgs = config 4 (24 contigs with GRCh38 length ratios, -g 1 -M, assigned largest-first to the GPUs: strong scaling)")
    ap.add_argument("--ploidy", type=int, default=None, help="-p (default 2; 4 under --workload tetra)")
    ap.add_argument("--A", type=int, default=None, help="-A windows sampling factor (default 2; 4 under --workload tetra)")
    ap.add_argument("--wgs-scale", type=float, default=1.0 / 16, help="--workload wgs: contig lengths = GRCh38 primary lengths x this (1.0 = 3.09 Gb)")
    a = ap.parse_args()
    if a.workload == "tetra":
        a.depth = 100.0 if a.depth == 30.0 else a.depth
        a.ploidy = a.ploidy or 4
        a.A = a.A or 4
    a.ploidy = a.ploidy or 2
    a.A = a.A or 2
    return a


def workload_name(a):
    if a.workload == "tetra":
        return (f"config 5: synthetic {a.depth:g}x paired-end 2x150 contig of {a.mb:g} Mb per GPU, -p {a.ploidy} -A {a.A} -M (read-depth window sweep at "
                f"{a.A} offsets per 10 kb frame, tetraploid thresholds), same evidence classes as config 3")
    return (f"config 3: synthetic {a.depth:g}x paired-end 2x150 contig of {a.mb:g} Mb per GPU (chr20-sized) with every evidence class "
            f"(soft/hard clips, SA tags, small indels, discordant pairs of all orientations, planted DEL/DUP/INV/CTX/INS clusters, "
            f"copy-number segments, 3 % low MAPQ, 5 % PCR duplicates), -M, SNV/indel/SV gates + read-depth CNV")


def params_for_bench(a=None):
    from grom_b200.params import Params
    kw = {}
    if a is not None:
        kw = dict(ploidy=a.ploidy, windows_sampling_factor=a.A)
        if a.workload == "wgs":
            kw["gender"] = 1
    return Params.default(insert_mean=400, insert_min=170, insert_max=520, lseq=150, rmdup=1, **kw)


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def window(self, t0, t1):
        """keep only the samples taken inside [t0, t1] (the timed regions)"""
        self.rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r[1:] for r in self.rows]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU baseline
HOT_TIMERS = (0, 1, 2, 3, 4, 7, 8)      # window management, -M, evidence, per-position scan, CNV depth, CNV pre-statistics, detect_del_dup (src/GROM.c:52-65, 18210-18220)


def bench_config(a):
    """The `config` object of the JSON line: the workload only (identical in both arms; run details live under other keys)."""
    return {"workload": workload_name(a), "contig_len": int(a.mb * 1e6), "depth": a.depth, "flags": "-M (duplicate filter on), defaults otherwise" if a.workload != "tetra" else f"-M -p {a.ploidy} -A {a.A}",
            "generator": ("tools/workloads.py chr20_spec" if not a.simple else "simple (single-M reads)"),
            "step": "evidence + SNV/indel/SV scan (gromgpu_chr_run) + read-depth CNV path (gromgpu_chr_cnv, incl. its host parts)",
            "partition": "one contig per GPU, no data-path collective",
            "l2": "inputs (>3 GB reads + 6.7 GB arrays per step) exceed the 126 MB L2; no flush needed"}


def cpu_reference_run(a, steps: int, warmup: int):
    """Time the reference's own implementation (oracle/_ref/GROM_ref_timing = reference src/GROM.c built with its own -DDO_TIMING
    instrumentation over a zlib-only samtools shim) on a bounded sample of the bench workload: P = cores/2 contigs of the config-3
    generator (+ a dummy last contig, which -P >= 2 silently skips, reference src/GROM.c:20999) processed by `-M -P P`, 2 threads per
    process (src/GROM.c:575).  Returns the whole-program figure (wall clock) and the hot-path-only one (sum of the reference's stage
    timers 0-4, 7, 8 over its children).  The sample is generated and written by a child process: this process maps no library of the repo."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    exe = os.path.join(ref_dir, "GROM_ref_timing")
    if not os.path.exists(exe):
        exe = os.path.join(ref_dir, "GROM_ref")
    if not os.path.exists(exe):
        return None
    cores = os.cpu_count() or 2
    nproc = max(1, min(cores // 2, 32))
    tmp = tempfile.mkdtemp(prefix="grom_cpu_")
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_sample.py"), "--out", tmp, "--contigs", str(nproc), "--mb", str(a.cpu_sample_mb),
                            "--depth", str(a.depth)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise RuntimeError("tools/make_sample.py failed: " + r.stderr[-400:])
        meta = json.loads(r.stdout.strip().splitlines()[-1])
        bases = meta["aligned_bases"]
        # the binary looks for its probability tables beside itself (src/GROM.c:21333, 21527): run a copy from the scratch directory
        run_exe = os.path.join(tmp, os.path.basename(exe))
        shutil.copy(exe, run_exe)
        for f in os.listdir(ref_dir):
            if f.endswith(".txt"):
                os.symlink(os.path.join(ref_dir, f), os.path.join(tmp, f))
        hz = None
        tsc = os.path.join(ref_dir, "tsc_hz")
        if os.path.exists(tsc):
            hz = float(subprocess.run([tsc], stdout=subprocess.PIPE, text=True).stdout.strip() or 0) or None
        times, hot = [], []
        extra = ["-p", str(a.ploidy), "-A", str(a.A)] if a.workload == "tetra" else (["-g", "1"] if a.workload == "wgs" else [])
        for it in range(warmup + steps):
            for ext in (".mean", ".info"):
                for q in (meta["bam"] + ext, meta["fasta"] + ext):
                    if os.path.exists(q):
                        os.remove(q)
            t0 = time.perf_counter()
            pr = subprocess.run([run_exe, "-i", meta["bam"], "-r", meta["fasta"], "-o", os.path.join(tmp, "out.vcf"), "-M", "-P", str(nproc)] + extra,
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            dt = time.perf_counter() - t0
            if pr.returncode != 0:
                raise RuntimeError(f"{os.path.basename(exe)} exited {pr.returncode}")
            if it >= warmup:
                times.append(dt)
                cyc = 0
                for line in pr.stdout.splitlines():
                    w = line.split()
                    if len(w) == 3 and w[0] == "timer" and w[1].isdigit() and int(w[1]) in HOT_TIMERS:
                        cyc += int(w[2])
                hot.append(cyc)
        sec = float(np.mean(times))
        out = {"value": bases / sec, "unit": UNIT, "cores": 2 * nproc, "kind": "reference",
               "sample": f"{os.path.basename(exe)} -M {' '.join(extra)} -P {nproc} on {nproc} contigs x {a.cpu_sample_mb:g} Mb of the config-3 generator at {a.depth:g}x "
                         f"({bases / 1e6:.0f} M aligned bases; whole program incl. two BAM decode passes, tables, VCF text), {sec:.2f} s wall, host has {cores} cores",
               "seconds": sec, "aligned_bases": bases}
        if hz and hot and hot[0] > 0:
            cpu_s = float(np.mean(hot)) / hz                   # CPU-seconds inside the hot path, all children together
            out["hot_path"] = {"cpu_seconds": cpu_s, "bases_per_s_per_process": bases / cpu_s, "bases_per_s_all_processes": bases / (cpu_s / nproc),
                               "processes": nproc, "timers": list(HOT_TIMERS), "tsc_hz": hz,
                               "what": "sum of the reference's own -DDO_TIMING stage timers (rdtsc) over its -P children: window management, -M, evidence "
                                       "accumulation, per-position scan, CNV depth, CNV pre-statistics, detect_del_dup; no BAM decode, FASTA load or VCF text"}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(a, max(1, min(a.steps, 5)), min(a.warmup, 1))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/GROM_ref was not built (needs /root/reference at build time)"}))
        return 0
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": bench_config(a),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "hot_path") if k in r},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "each timed step = one whole run of the reference on the bounded sample (at most 5 are run, whatever --steps says)"}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- B200 arm
def pin_batch(batch):
    """Copy every array of the batch (and of its transport-compact forms, if attached) into pinned host memory; returns
    (batch view over pinned memory, keepalive, bytes the library copies host -> device per push)."""
    import torch
    from grom_b200.reads import _DTYPES, ReadBatch

    keep = []

    def pin(src):
        t = torch.from_numpy(src).pin_memory() if src.size else torch.from_numpy(src)
        keep.append(t)
        return t.numpy()

    out = ReadBatch(tid=batch.tid, **{k: pin(getattr(batch, k)) for k in _DTYPES})
    out.layout_flags = batch.layout_flags
    if batch.qual4 is not None:
        out.qual4, out.qual_lut = pin(batch.qual4), batch.qual_lut
    if batch.qual2 is not None:
        out.qual2, out.qual_lut = pin(batch.qual2), batch.qual_lut
    if batch.seq2 is not None:
        out.seq2, out.seq_exc_slot, out.seq_exc_code = pin(batch.seq2), pin(batch.seq_exc_slot), pin(batch.seq_exc_code)
    if batch.sa_index is not None:
        out.sa_index, out.sa_sparse = pin(batch.sa_index), {k: pin(v) for k, v in batch.sa_sparse.items()}
    return out, keep, out.transport_bytes()


def _cpulist(text):
    out = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def plan_cores(allowed, gpu_nodes, node_cpus, local):
    """Cores of local rank `local`: `allowed` = cores this process may use, gpu_nodes[i] = NUMA node of local rank i's GPU (-1 unknown),
    node_cpus[n] = cores of NUMA node n.  Ranks whose GPUs share a node split that node's allowed cores evenly; when a node offers fewer than
    two allowed cores per rank (or is unknown) all allowed cores are split by local rank instead.  Returns (cores, node, ranks sharing)."""
    allowed = sorted(allowed)
    mine = gpu_nodes[local]
    pool, peers = allowed, list(range(len(gpu_nodes)))
    if mine >= 0:
        on_node = sorted(set(node_cpus.get(mine, ())) & set(allowed))
        sharing = [i for i in range(len(gpu_nodes)) if gpu_nodes[i] == mine]
        # every rank must come to the same decision, so the rule may only depend on its own node's numbers when all nodes pass it
        ok = all(len(set(node_cpus.get(n, ())) & set(allowed)) >= 2 * sum(1 for x in gpu_nodes if x == n) for n in set(gpu_nodes) if n >= 0) and all(n >= 0 for n in gpu_nodes)
        if ok:
            pool, peers = on_node, sharing
    k, n = peers.index(local), len(peers)
    return pool[len(pool) * k // n:len(pool) * (k + 1) // n], mine, n


def bind_rank_to_cores(local: int, local_world: int):
    """N > 1 ranks on one node: give every rank its own cores, on the NUMA node its GPU hangs off where the process is allowed there.
    The ranks share nothing on the data path but the host: without this the lane / OpenMP / CNV host threads of all ranks migrate over all
    cores, and the pinned upload buffers (first touched by this process) may sit on the other socket from the GPU's PCIe root.  Returns a
    description for the JSON line, or None (one rank, no sysfs, a refused affinity call: nothing is changed then)."""
    if local_world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch

        def gpu_node(i):
            pr = torch.cuda.get_device_properties(i % max(1, torch.cuda.device_count()))
            bus = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{getattr(pr, 'pci_device_id', 0):02x}.0"
            try:
                return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            except OSError:
                return -1
        nodes = [gpu_node(i) for i in range(local_world)]
        node_cpus = {}
        for n in set(nodes):
            if n >= 0:
                try:
                    node_cpus[n] = _cpulist(open(f"/sys/devices/system/node/node{n}/cpulist").read())
                except OSError:
                    node_cpus[n] = set()
        share, mine, n = plan_cores(os.sched_getaffinity(0), nodes, node_cpus, local)
        if len(share) < 2:
            return None
        os.sched_setaffinity(0, share)
        return {"cpus": f"{share[0]}-{share[-1]}" if share == list(range(share[0], share[-1] + 1)) else ",".join(map(str, share)),
                "n_cpus": len(share), "gpu_numa_node": mine, "ranks_sharing_the_pool": n}
    except Exception as e:                                   # binding is an optimisation, never a reason to fail the run
        return {"error": str(e)[:120]}


def main_b200(a):
    import torch
    import torch.distributed as dist
    from grom_b200 import gpu, hostlib
    from tools import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    # host threads per gromgpu_chr_cnv call: the node's cores shared by the ranks of the node (the contigs in flight of one rank are rarely in
    # their host stages at the same time)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    binding = None if a.no_bind else bind_rank_to_cores(local, local_world)
    n_mine = binding.get("n_cpus") if binding else None
    os.environ.setdefault("GROMGPU_HOST_THREADS", str(n_mine or max(2, (os.cpu_count() or 2) // max(1, local_world))))
    if n_mine:
        os.environ.setdefault("OMP_NUM_THREADS", str(n_mine))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prm = params_for_bench(a)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    # synthetic chr20-sized contig of this rank (rank-specific seed); inputs (>3 GB) far exceed the 126 MB L2
    P = int(a.mb * 1e6)
    from tools import workloads
    if a.simple:
        spec = synth.SynthSpec(contigs=[(f"chr{20 + rank}", P)], depth=a.depth, seed=20 + rank, simple=True, dup_frac=0.05, names=False, simple_disc_frac=0.01)
    else:
        spec = workloads.chr20_spec(mb=a.mb, depth=a.depth, seed=20 + rank, name=f"chr{20 + rank}", cnv_per_mb=a.cnv_per_mb)
    t0 = time.time()
    c = synth.simulate(spec)[0]
    gen_s = time.time() - t0
    # the batch in the form the host batcher hands over: canonical arrays + the transport-compact forms the data allow
    # (repack_canonical = the layout the batcher produces natively: every read's bases start on a 32-slot boundary, offsets are running sums)
    pinned, keep, read_bytes = pin_batch(c.batch if a.canonical_upload else c.batch.repack_canonical().compact())
    fasta_pinned = torch.from_numpy(c.chars).pin_memory()
    fasta_np = fasta_pinned.numpy()

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        gpu.init(local, hez, mq, prm)
        gpu.set_stream(stream.cuda_stream)
        ch = gpu.Chromosome(c.batch.tid, fasta_np)
        ch.push_reads(pinned)
        # ---- resident-input timing (value): W warm-up + K timed steps
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(max(3, a.warmup)):
            ch.run(); ch.cnv()
        barrier()
        t_region0 = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = {k: 0.0 for k in ("ms_total", "ms_clear", "ms_gc", "ms_dup", "ms_prep", "ms_index", "ms_sv", "ms_rdscan", "ms_pileup")}
        launches = 0
        cnv_ms = {"ms_cnv_device": 0.0, "ms_cnv_host": 0.0}
        ev0.record(stream)
        for _ in range(a.steps):
            ch.run()
            cn = ch.cnv()
            s = ch.stats()
            for k in per:
                per[k] += getattr(s, k)
            launches += s.launches + cn.launches
            cnv_ms["ms_cnv_device"] += cn.ms_device; cnv_ms["ms_cnv_host"] += cn.ms_host
        ev1.record(stream)
        barrier()
        ms_steps = ev0.elapsed_time(ev1)
        st = ch.stats()
        res = ch.result()
        # ---- end-to-end timing through the C ABI with host (pinned) buffers: reset + FASTA H2D + reads H2D + kernels + result D2H.
        # (a) one contig at a time; (b) the per-genome driver's shape: two contigs in flight on two streams / host threads, so the
        # upload of contig i+1 overlaps the kernels and the host part of contig i (every step still copies its own inputs and results)
        d2h_bytes = len(res.snv) * 128 + len(res.ins) * 104 + len(res.del_ev) * 48 + len(res.sv_ev) * 64 + 64 + cn.d2h_bytes + len(cn.calls) * 56
        for _ in range(2):
            ch.reset(fasta_np); ch.push_reads(pinned); ch.finish(); ch.cnv()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            ch.reset(fasta_np)
            ch.push_reads(pinned)
            r2 = ch.finish()
            cn2 = ch.cnv()
        e1.record(stream)
        barrier()
        ms_e2e_seq = e0.elapsed_time(e1)
        assert len(r2.snv) == len(res.snv) and len(cn2.calls) == len(cn.calls)
        import threading
        # contigs in flight: as many as asked for, as long as their handles fit beside the first one (same admission rule as the genome drivers)
        need = gpu.chr_bytes_estimate(P, c.batch.n_reads, c.batch.n_base_slots)
        n_lanes = max(1, min(a.lanes, 1 + int(0.9 * gpu.device_free_bytes() // max(1, need))))
        lanes = [(ch, stream)]
        for _ in range(n_lanes - 1):
            st_x = torch.cuda.Stream()
            gpu.set_stream(st_x.cuda_stream)
            lanes.append((gpu.Chromosome(c.batch.tid, fasta_np), st_x))
        gpu.set_stream(stream.cuda_stream)
        done_ev = [torch.cuda.Event(enable_timing=True) for _ in lanes]
        counts = [None] * n_lanes
        lane_log = []                                           # (lane, ms waiting for the bus incl. reset, ms upload, ms finish, ms cnv, cnv device, cnv host) per step
        bus = threading.Lock()                                  # one upload at a time: the PCIe link is the shared resource
        todo = [0]
        todo_lock = threading.Lock()

        def lane(k, n_steps, record):
            torch.cuda.set_device(local)
            h, st_k = lanes[k]
            while True:
                with todo_lock:                                 # steps are handed out as lanes become free
                    if todo[0] >= n_steps:
                        break
                    todo[0] += 1
                t_a = time.perf_counter()
                h.reset(fasta_np)
                with bus:
                    t_b = time.perf_counter()
                    h.push_reads(pinned); st_k.synchronize()
                    t_c = time.perf_counter()
                rr = h.finish(); t_d = time.perf_counter(); cc = h.cnv()
                lane_log.append((k, round((t_b - t_a) * 1e3, 1), round((t_c - t_b) * 1e3, 1), round((t_d - t_c) * 1e3, 1), round((time.perf_counter() - t_d) * 1e3, 1),
                                 round(cc.ms_device, 1), round(cc.ms_host, 1)))
                counts[k] = len(rr.snv) + len(cc.calls)
            if record:
                done_ev[k].record(st_k)

        def run_pipelined(n_steps, record):
            todo[0] = 0
            th = [threading.Thread(target=lane, args=(k, n_steps, record)) for k in range(n_lanes)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        for _ in range(2):
            run_pipelined(n_lanes, False)                      # warm every lane (first call allocates)
        barrier()
        p0 = torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _, st_x in lanes[1:]:
            st_x.wait_stream(stream)
        run_pipelined(a.steps, True)
        barrier()
        ms_e2e = max(p0.elapsed_time(e) for e in done_ev)
        assert all(x is None or x == len(res.snv) + len(cn.calls) for x in counts) and any(x is not None for x in counts)
        if rank == 0:
            sys.stderr.write("e2e lanes (lane, wait, upload, finish, cnv, cnv device, cnv host ms): %s\n" % lane_log[-min(8, len(lane_log)):])
        for h_x, _ in lanes[1:]:
            h_x.close()
        if rank == 0:
            sampler.window(t_region0, time.time())
        clocks = sampler.stop() if rank == 0 else None
        ch.close()

    bases = int(st.aligned_bases)
    tmax = torch.tensor([ms_steps, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([bases], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_steps_max, ms_e2e_max = float(tmax[0]), float(tmax[1])
    total_bases = float(tot[0])

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # dominant kernel = pileup (+ fused SNV gate): algorithmic bytes = read records consumed + the 26 int32 arrays it writes
        # + the 2 it reads for the gate + the FASTA char
        alg_bytes = st.bytes_reads + (4 * (N_ARRAYS_PILEUP + N_ARRAYS_PILEUP_READ) + 1) * P
        ms_pile = per["ms_pileup"] / a.steps
        achieved = alg_bytes / (ms_pile * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "pileup_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch_scaled_to_mb", {}).get(f"{a.mb:g}") or tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # ---- after the timed loops: the same generator at --parity-mb through the same library, every array / candidate / call against the oracle
        parity, parity_counts, decode = "skipped", None, None
        if a.parity_mb > 0:
            t_p = time.time()
            spec_p = workloads.chr20_spec(mb=a.parity_mb, depth=a.depth, seed=2020, name="chr20p", names=True, cnv_per_mb=max(a.cnv_per_mb, 0.5))
            cp = synth.simulate(spec_p)
            from tools import parity as parity_mod
            try:
                parity_counts = parity_mod.compare_gpu_oracle(prm, cp[0], hez, mq, device=local)
                parity = "ok"
            except AssertionError as e:
                parity = f"FAILED: {e}"
            except Exception as e:                       # the checker itself broke (not a difference): say so, keep the measured line
                parity = f"ERROR: {type(e).__name__}: {str(e)[:200]}"
            parity_counts = dict(parity_counts or {}, seconds=round(time.time() - t_p, 1), contig_mb=a.parity_mb,
                                 what="all 56 per-position arrays, -M flags, 10 breakpoint clusters, SNV / indel / SV gate records, CNV mask + z + window table + calls, bit-exact vs oracle/")
            # ---- host batcher (BAM decode) throughput on the same data, reported separately (north star)
            tmpd = tempfile.mkdtemp(prefix="grom_dec_")
            try:
                fa_p, bam_p = synth.write_dataset(os.path.join(tmpd, "p"), cp)
                nthr = int(os.environ.get("OMP_NUM_THREADS", "0")) or (os.cpu_count() or 1)
                best = None
                for _ in range(3):
                    t_d = time.perf_counter()
                    with hostlib.Bam(bam_p) as bf:
                        bt = bf.read_target_owned(0)          # the product call (the batch stays in the batcher's memory, as in tools/grom_b200.c)
                    dt = time.perf_counter() - t_d
                    bt.free()
                    best = dt if best is None else min(best, dt)
                ab = cp[0].batch.aligned_bases()
                decode = {"bases_per_s": ab / best, "threads": nthr, "reads_per_s": bt.n_reads / best, "bam_bytes": os.path.getsize(bam_p),
                          "what": f"gromhost_bam_open + gromhost_bam_read_target (BGZF inflate + record parse + SA pre-parse -> packed SoA batch + transport-compact forms) on a {a.parity_mb:g} Mb / {a.depth:g}x BAM, best of 3"}
            except Exception as e:                       # the separately reported host share must not take the line down with it
                decode = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            finally:
                shutil.rmtree(tmpd, ignore_errors=True)
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            r = cpu_reference_run(a, 1, 0)
            if r:
                cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "hot_path") if k in r}
        line = {
            "metric": METRIC, "value": total_bases * a.steps / (ms_steps_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms_steps_max / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": bench_config(a),
            "run": {"reads_per_gpu": int(st.n_reads), "aligned_bases_per_gpu": bases, "host_gen_s": round(gen_s, 1), "host_cores_rank0": binding},
            "e2e": {"value": total_bases * a.steps / (ms_e2e_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(read_bytes + P),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e_max / a.steps,
                    "upload_form": ("canonical arrays" if not pinned.layout_flags else "transport-compact (include/grom_reads.h GROM_LAYOUT_*): "
                                    + ", ".join(n for bit, n in ((1, "offsets derived on the device"), (2, "4-bit dictionary qualities"), (16, "2-bit dictionary qualities"), (8, "2-bit bases + exception list"), (4, "sparse SA fields")) if pinned.layout_flags & bit)),
                    "mode": f"{n_lanes} contigs in flight (one stream / host thread each, uploads serialised): the upload of step i+1 overlaps the kernels and host part of step i",
                    "one_at_a_time": {"value": bases * a.steps / (ms_e2e_seq * 1e-3), "ms_per_step": ms_e2e_seq / a.steps}},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_pileup", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes), "ms_per_launch": ms_pile},
            "kernels_ms_per_step": {**{k: v / a.steps for k, v in per.items()}, **{k: v / a.steps for k, v in cnv_ms.items()}},
            "cpu_baseline": cpu,
            "parity_check": parity, "parity_detail": parity_counts, "decode": decode,
            "clocks": clocks,
            "results": {"snv_candidates": int(len(res.snv)), "dups": int(st.n_dups), "applied_reads": int(st.n_applied),
                        "sv_items": int(st.n_sv_items), "small_ins": int(len(res.ins)),
                        "small_del_events": int(len(res.del_ev)), "cnv_calls": int(len(cn.calls))},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    # stdout carries exactly one JSON line: everything libraries print there (e.g. NCCL's version banner) is sent to stderr
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*a, **k):                                        # noqa: A001 - the two arms print their JSON line through this
        _buf.append(" ".join(str(x) for x in a))
    if args.workload == "wgs" and args.impl != "reference":
        from tools import wgs_bench
        rc = wgs_bench.main_wgs(args, print)
    else:
        rc = main_reference(args) if args.impl == "reference" else main_b200(args)
    sys.stdout.flush()
    os.dup2(_real_stdout, 1)
    for line in _buf:
        _print(line, flush=True)
    sys.exit(rc)
