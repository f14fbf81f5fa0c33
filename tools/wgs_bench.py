"""bench.py --workload wgs: config 4 of BASELINE.json -- a synthetic 30x whole genome (24 contigs with the GRCh38 primary-assembly
length ratios, -g 1 -M), contigs assigned to the GPUs largest-first like the reference's -P scheduler (reference src/GROM.c:22318-22336,
549-599; grom_b200/partition.py), no data-path collective.  STRONG scaling: the genome is fixed, the ranks share it.

One step = one pass over the whole genome through the C ABI with HOST (pinned) buffers, exactly the per-contig flow of the product
drivers (grom_b200/pipeline.py, tools/grom_b200.c): gromgpu_chr_begin_on / gromgpu_chr_rebind (FASTA host -> device) -> gromgpu_push_reads (packed reads
host -> device) -> gromgpu_chr_run + gromgpu_chr_result (evidence, SNV / indel / SV gates, candidates device -> host) -> gromgpu_chr_cnv
(read-depth path incl. its host parts), up to --lanes contigs in flight per GPU, each lane on ONE handle that is begun for its largest contig and
rebound (gromgpu_chr_rebind: buffers kept, arrays zeroed) for the others.  Every rank generates only its own
contigs (seed = 38 + contig index, so a contig's reads do not depend on the number of ranks), which is also why the per-contig record
text must be identical at every N: its digest is part of the JSON line (`genome_digest`) and tests/test_gpu_wgs.py compares 1 rank
against 2.  BAM decode is not in the timed region (its throughput is the `decode` object of the default bench line).

The default scale is 1/16 (193 Mb, 38.6 M reads): the generator is numpy on the host cores and a full-size genome would take
~20 minutes to synthesise; pass --wgs-scale 1 for GRCh38 lengths (chr1 = 109 GB of device arrays, fits the 180 GB of one B200).
Nothing here is on the product path.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259), ("chr6", 170805979),
          ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
          ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285),
          ("chr19", 58617616), ("chr20", 64444167), ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


def genome(scale: float):
    return [(n, max(200_000, int(l * scale))) for n, l in GRCH38]


def _make_contig(job):
    """(tid, name, length, depth, cnv_per_mb) -> (tid, FASTA characters, packed read batch in the transport-compact forms)"""
    tid, name, length, depth, cnv_per_mb = job
    from tools import synth, workloads
    spec = workloads.chr20_spec(mb=length / 1e6, depth=depth, seed=38 + tid, name=name, cnv_per_mb=cnv_per_mb, dummy_len=100_000)
    spec.contigs = [(name, length), ("chrzz", 100_000)]
    c = synth.simulate(spec)[0]
    return tid, c.chars, c.batch.repack_canonical().compact()


def main_wgs(a, emit) -> int:
    from grom_b200.partition import assign_contigs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    contigs = genome(a.wgs_scale)
    lens = [l for _, l in contigs]
    plan = assign_contigs(lens, world)
    mine = plan[rank]                                               # largest first
    # ---- this rank's contigs, generated on the host cores before any CUDA state exists (worker processes are forked)
    t0 = time.time()
    jobs = [(t, contigs[t][0], contigs[t][1], a.depth, a.cnv_per_mb) for t in mine]
    nproc = max(1, min(len(jobs), (os.cpu_count() or 2) // max(1, world) - 1))
    if nproc > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(nproc) as pool:
            made = pool.map(_make_contig, jobs, chunksize=1)
    else:
        made = [_make_contig(j) for j in jobs]
    gen_s = time.time() - t0

    import torch
    import torch.distributed as dist
    from bench import ClockSampler, params_for_bench, pin_batch, cpu_reference_run, METRIC, UNIT
    from grom_b200 import gpu, hostlib
    from grom_b200.pipeline import _InFlight
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    from bench import bind_rank_to_cores
    binding = None if getattr(a, "no_bind", False) else bind_rank_to_cores(local, local_world)      # own cores per rank, near its GPU
    n_mine = binding.get("n_cpus") if binding else None
    os.environ.setdefault("GROMGPU_HOST_THREADS", str(n_mine or max(2, (os.cpu_count() or 2) // max(1, local_world))))
    ndev = torch.cuda.device_count()
    dev = local % ndev                                              # (tests run two ranks on one GPU)
    torch.cuda.set_device(dev)
    if world > 1:
        backend = os.environ.get("GROM_DIST_BACKEND") or ("nccl" if ndev >= int(os.environ.get("LOCAL_WORLD_SIZE", world)) else "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        else:
            dist.init_process_group("gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prm = params_for_bench(a)
    hez, mq = hostlib.tables(None, prm.min_mapq)
    gpu.init(dev, hez, mq, prm)
    items, keep, h2d = {}, [], 0
    for tid, chars, batch in made:
        pb, k, nbytes = pin_batch(batch)
        fa = torch.from_numpy(chars).pin_memory()
        keep += k + [fa]
        items[tid] = (fa.numpy(), pb, batch.n_reads, batch.n_base_slots, int(batch.tid))     # the generator numbers every contig 0 and its mates' dummy contig 1
        h2d += nbytes + len(chars)
    del made
    n_lanes = max(1, min(a.lanes, len(mine)))
    streams = [torch.cuda.Stream() for _ in range(n_lanes)]
    main_stream = torch.cuda.current_stream()
    inflight = _InFlight(int(0.9 * gpu.device_free_bytes()))
    slots = [[None, 0] for _ in range(n_lanes)]
    lane_log = []                                                    # (lane, contig, ms bus + rebind + upload, ms run + result, ms cnv, cnv device, cnv host, cnv D2H bytes)

    def one_pass(collect):
        """all of this rank's contigs, largest first, n_lanes in flight; returns (ms on the device, per-contig info)"""
        work = list(mine)
        pick, bus, errors = threading.Lock(), threading.Lock(), []
        info = {}
        done = [torch.cuda.Event(enable_timing=True) for _ in range(n_lanes)]
        p0 = torch.cuda.Event(enable_timing=True)
        p0.record(main_stream)
        for s in streams:
            s.wait_stream(main_stream)

        def lane(k):
            torch.cuda.set_device(dev)
            try:
                # lane k takes contigs k, k + n_lanes, ... of the largest-first list: its first one is the largest it ever sees, in every
                # pass, so its handle never has to grow after the first pass (a dynamic queue would hand lanes different contigs each pass)
                for t in work[k::n_lanes]:
                    if errors:
                        break
                    chars, pb, n_reads, n_slots, btid = items[t]
                    # one handle per lane, begun for the largest contig the lane sees and rebound for the others (the product drivers'
                    # flow: grom_b200/pipeline.py, tools/grom_b200.c); it stays alive across passes like it does across contigs
                    slot = slots[k]
                    t_a = time.perf_counter()
                    with bus:                                        # one upload at a time: the PCIe link is the shared resource
                        if slot[0] is None or not slot[0].rebind(btid, chars):
                            if slot[0] is not None:
                                slot[0].close(); inflight.release(slot[1]); slot[0] = None
                            need = gpu.chr_bytes_estimate(len(chars), n_reads, n_slots)
                            inflight.acquire(need)
                            slot[1] = need
                            slot[0] = gpu.Chromosome(btid, chars, stream=streams[k].cuda_stream)
                        ch = slot[0]
                        ch.push_reads(pb); ch.sync()
                    t_b = time.perf_counter()
                    res = ch.finish()
                    t_c = time.perf_counter()
                    cn = ch.cnv(params=prm)
                    t_d = time.perf_counter()
                    st = ch.stats()
                    lane_log.append((k, t, round((t_b - t_a) * 1e3, 1), round((t_c - t_b) * 1e3, 1), round((t_d - t_c) * 1e3, 1), round(cn.ms_device, 1), round(cn.ms_host, 1), int(cn.d2h_bytes)))
                    d = {"bases": int(st.aligned_bases), "reads": int(st.n_reads), "ms_pileup": st.ms_pileup, "ms_run": st.ms_total, "ms_cnv_device": cn.ms_device,
                         "ms_cnv_host": cn.ms_host, "launches": int(st.launches + cn.launches), "bytes_reads": int(st.bytes_reads),
                         "d2h": int(len(res.snv) * 128 + len(res.ins) * 104 + len(res.del_ev) * 48 + len(res.sv_ev) * 64 + 64 + cn.d2h_bytes + len(cn.calls) * 56),
                         "records": None}
                    if collect:
                        text = hostlib.vcf_contig(prm, contigs[t][0].lower(), chars, res.snv, res.snv_ave_rd, res.ins, res.del_ev, res.sv_ev, cn.calls)
                        d["sha1"] = hashlib.sha1(text.encode()).hexdigest(); d["records"] = text.count("\n")
                        d["snv"], d["cnv_calls"], d["sv_events"] = int(len(res.snv)), int(len(cn.calls)), int(len(res.sv_ev))
                    info[t] = d
            except BaseException as e:                                # surfaced below
                errors.append(e)
            done[k].record(streams[k])

        th = [threading.Thread(target=lane, args=(k,)) for k in range(n_lanes)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errors:
            raise errors[0]
        torch.cuda.synchronize()
        return max(p0.elapsed_time(e) for e in done), info

    sampler = ClockSampler(dev)
    warm = max(3, a.warmup)
    for _ in range(warm):
        one_pass(False)
    barrier()
    if rank == 0:
        sampler.start()
    t_region0 = time.time()
    ms, infos = 0.0, None
    barrier()
    w0 = time.perf_counter()
    for _ in range(a.steps):
        m, infos = one_pass(False)
        ms += m
    wall_ms = (time.perf_counter() - w0) * 1e3
    barrier()
    if rank == 0:
        sampler.window(t_region0, time.time())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        sys.stderr.write("wgs lanes (lane, contig, upload, run, cnv, cnv device, cnv host ms, cnv d2h): %s\n" % lane_log[-min(24, len(lane_log)):])
    _, final = one_pass(True)                                       # untimed: the record text of every contig, for the digest
    for sl in slots:
        if sl[0] is not None:
            sl[0].close()

    tmax = torch.tensor([ms, wall_ms], dtype=torch.float64)
    mine_row = {"rank": rank, "device": dev, "contigs": [contigs[t][0] for t in mine], "positions": int(sum(lens[t] for t in mine)),
                "bases": int(sum(infos[t]["bases"] for t in mine)), "h2d": int(h2d), "ms_per_step": ms / a.steps, "gen_s": round(gen_s, 1), "per_contig": {str(t): final[t] for t in mine}}
    rows = [mine_row]
    if world > 1:
        tm = tmax.cuda() if dist.get_backend() == "nccl" else tmax
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tmax = tm.cpu()
        rows = [None] * world
        dist.all_gather_object(rows, mine_row)
    if rank == 0:
        ms_max = float(tmax[0])
        total_bases = sum(r["bases"] for r in rows)
        per = {}
        for r in rows:
            per.update({int(k): v for k, v in r["per_contig"].items()})
        digest = hashlib.sha1("".join(per[t]["sha1"] for t in sorted(per)).encode()).hexdigest()
        loads = [r["positions"] for r in rows]
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = float(peaks.get("hbm_gbs", 6650.0))
        P_all = sum(lens)
        alg = sum(v["bytes_reads"] for v in per.values()) + (4 * 28 + 1) * P_all
        ms_pile = sum(v["ms_pileup"] for v in per.values())
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            r = cpu_reference_run(a, 1, 0)
            if r:
                cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "hot_path") if k in r}
        val = total_bases * a.steps / (ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms_max / a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"config 4: synthetic {a.depth:g}x paired-end 2x150 whole genome, 24 contigs with the GRCh38 primary-assembly lengths x {a.wgs_scale:g} "
                                   f"({P_all / 1e6:.0f} Mb), -g 1 -M, every evidence class of config 3, contigs assigned to the GPUs largest-first (the reference's -P policy)",
                       "genome_positions": int(P_all), "depth": a.depth, "flags": "-g 1 -M", "generator": "tools/wgs_bench.py (tools/workloads.py chr20_spec per contig, seed 38 + contig index)",
                       "step": "one pass over the genome: per contig gromgpu_chr_rebind (FASTA host -> device, arrays zeroed) + packed reads host -> device, gromgpu_chr_run / result, gromgpu_chr_cnv; one handle per lane, begun in the first pass",
                       "partition": "largest-first greedy over contig lengths, no data-path collective", "l2": "every contig's inputs exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": int(sum(r["h2d"] for r in rows)), "d2h_bytes_per_step": int(sum(v["d2h"] for v in per.values())),
                    "ms_per_step": ms_max / a.steps, "wall_ms_per_step": float(tmax[1]) / a.steps,
                    "note": "this workload is only measured end to end (pinned host buffers, uploads inside the timed region); `value` repeats it"},
            "gpu_launches": int(sum(v["launches"] for v in per.values()) * a.steps),
            "roofline": {"kernel": "k_pileup", "bound": "hbm", "achieved": alg / (ms_pile * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms_pile * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "algorithmic_bytes_per_launch": int(alg), "ms_per_launch": ms_pile, "note": "summed over the 24 launches of one pass"},
            "cpu_baseline": cpu,
            "lpt": {"loads_positions": loads, "imbalance": max(loads) / (sum(loads) / len(loads)), "ms_per_step_per_rank": [round(r["ms_per_step"], 2) for r in rows],
                    "contigs_per_rank": [r["contigs"] for r in rows]},
            "device_ms_per_step": {"run": sum(v["ms_run"] for v in per.values()), "cnv_device": sum(v["ms_cnv_device"] for v in per.values()),
                                   "cnv_host": sum(v["ms_cnv_host"] for v in per.values())},
            "genome_digest": digest, "records": int(sum(v["records"] for v in per.values())),
            "results": {"snv_candidates": int(sum(v["snv"] for v in per.values())), "cnv_calls": int(sum(v["cnv_calls"] for v in per.values())),
                        "sv_events": int(sum(v["sv_events"] for v in per.values())), "reads": int(sum(v["reads"] for v in per.values()))},
            "run": {"host_gen_s": [r["gen_s"] for r in rows], "lanes": n_lanes, "host_cores_rank0": binding},
            "clocks": clocks,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0
