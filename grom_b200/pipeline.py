"""Per-genome driver over the C ABI: the analogue of the reference's find_disc_svs loop (src/GROM.c:20900-21130).

    BAM  --gromhost_bam_read_target-->  packed batches  --gromhost_libstats-->  insert statistics (find_insert_mean)
    per contig:  gromgpu_chr_begin / push_reads / chr_finish / chr_cnv  -->  gromhost_vcf_contig  -->  record text

Contigs are processed in BAM header order (the order the reference concatenates per-contig outputs in, src/GROM.c:21121-21126);
with `ranks` > 1 they are assigned largest-first (`partition.assign_contigs`) and only this rank's share is processed.
No CPU fallback: `gpu.init` fails without a CUDA device.
"""
from __future__ import annotations

import gzip
import threading
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import gpu, hostlib
from .params import Params
from .partition import assign_contigs, skip_contig


def read_fasta(path: str) -> Dict[str, np.ndarray]:
    """FASTA -> {first word of the header: uint8 characters, case preserved} (the reference loader, src/GROM.c:21009-21045)."""
    op = gzip.open if path.endswith(".gz") else open
    out: Dict[str, np.ndarray] = {}
    name, parts = None, []
    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    out[name] = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
                name = line[1:].split()[0].decode(); parts = []
            else:
                parts.append(line.rstrip(b"\r\n"))
    if name is not None:
        out[name] = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    return out


class _LazyFasta:
    """name (lower case) -> characters, one contig at a time through libgromhost's FASTA reader (gromhost_fasta_*, the reference's line
    rules); gzip-compressed files go through the Python reader above, whole."""

    def __init__(self, path: str):
        self._all = None
        if path.endswith(".gz"):
            self._all = {k.lower(): v for k, v in read_fasta(path).items()}
        else:
            self._fa = hostlib.Fasta(path)
            self._lock = threading.Lock()

    def __contains__(self, name: str) -> bool:
        return (name in self._all) if self._all is not None else self._fa.find(name) >= 0

    def __getitem__(self, name: str) -> np.ndarray:
        if self._all is not None:
            return self._all[name]
        with self._lock:
            k = self._fa.find(name)
            if k < 0:
                raise KeyError(name)
            return self._fa.load(k)


class _InFlight:
    """Admission control for the contigs in flight on one GPU: a contig starts when a lane is free and the device memory its handle
    will hold (gromgpu_chr_bytes_estimate) fits beside the ones already running; a contig that fits nowhere runs alone."""

    def __init__(self, budget: int):
        self.budget, self.used, self.running = budget, 0, 0
        self.cv = threading.Condition()

    def acquire(self, need: int):
        with self.cv:
            while self.running and self.used + need > self.budget:
                self.cv.wait()
            self.used += need; self.running += 1

    def release(self, need: int):
        with self.cv:
            self.used -= need; self.running -= 1
            self.cv.notify_all()


def call_variants(bam_path: str, fasta_path: str, params: Optional[Params] = None, device: int = 0, rank: int = 0, ranks: int = 1,
                  table_dir: Optional[str] = None, ctx_out: Optional[Dict[int, np.ndarray]] = None, lanes: int = 3,
                  slice_reads: Optional[int] = None) -> Tuple[Dict[int, str], Params]:
    """Returns ({tid: record text of that contig}, the parameters incl. the library statistics measured from the BAM).  If `ctx_out` is
    given it receives {tid: translocation records of that contig} for `ctx_vcf_text` (the pairing needs the records of all contigs).
    Up to `lanes` contigs are in flight on the GPU (one host thread and one stream each; uploads take turns on the PCIe link), so the
    upload of one contig overlaps the kernels and the host stages of the others; results do not depend on `lanes`.
    `slice_reads` (needs an index with record counts): a contig's reads are decoded and pushed in consecutive pieces of about that many
    records instead of as one batch, which bounds the host memory of a lane by the piece (gromhost_bam_iter_*); results do not depend on it."""
    prm = params if params is not None else Params.default()
    fasta = _LazyFasta(fasta_path)          # contig characters are loaded (by the C library) when a lane takes the contig, not all up front
    with hostlib.Bam(bam_path) as bam:
        # library statistics first (find_insert_mean, src/GROM.c:1205-1318): one windowed pass over the file that ends when the sample is full
        st = bam.library_stats(prm.min_mapq)
        prm.insert_mean = max(st["insert_mean"], st["lseq"])            # src/GROM.c:22260
        prm.insert_min, prm.insert_max, prm.lseq = st["insert_min"], st["insert_max"], st["lseq"]
        prm.rd_min_mapq = prm.min_mapq                                    # src/GROM.c:22102
        hez, mq = hostlib.tables(table_dir, prm.min_mapq)
        gpu.init(device, hez, mq, prm)
        todo = [t for t, n in enumerate(bam.names) if n.lower() in fasta and not skip_contig(n, prm.gender)]
        # load of a contig = its records where the index counts them (coverage differs between contigs), else its length (the reference's -P order)
        weights = [float(bam.read_counts[t]) for t in todo] if bam.read_counts is not None else None
        mine = set(todo[i] for i in assign_contigs([bam.lens[t] for t in todo], ranks, weights)[rank])
        text: Dict[int, str] = {}
        work = [t for t in todo if t in mine]
        work.sort(key=lambda t: -bam.lens[t])                              # largest first, like the reference's -P scheduler (src/GROM.c:22318-22336)
        n_lanes = max(1, min(lanes, len(work)))
        inflight = _InFlight(int(0.9 * gpu.device_free_bytes()))
        bus, pick, errors = threading.Lock(), threading.Lock(), []
        names, lens, read_counts = list(bam.names), list(bam.lens), bam.read_counts

        def one_contig(lane_bam, t: int, stream: Optional[int], slot: list):
            """slot = [handle, reserved bytes] of this lane: the handle is begun for the lane's first contig (the largest it will see,
            the queue is sorted largest first) and rebound -- same device buffers -- for the rest (gromgpu_chr_rebind)"""
            name = names[t].lower()
            chars = fasta[name]
            if len(chars) != lens[t]:
                raise ValueError(f"{names[t]}: {len(chars)} bases in the FASTA, {lens[t]} in the BAM header")
            sliced = bool(slice_reads) and read_counts is not None
            if sliced:                                                      # totals from the index, pieces decoded one at a time below
                n_total = int(read_counts[t]); slots_total = n_total * ((prm.lseq + 31) // 32 * 32)
                batch = None
            else:
                batch = lane_bam.read_target_owned(t)                       # decoded just before it is pushed (the batcher's own memory goes to the CUDA library), dropped right after
                n_total, slots_total = batch.n_reads, batch.n_base_slots
            if slot[0] is None or not slot[0].rebind(t, chars):
                if slot[0] is not None:
                    slot[0].close(); inflight.release(slot[1]); slot[0] = None
                need = gpu.chr_bytes_estimate(len(chars), n_total, slots_total)
                inflight.acquire(need)
                slot[1] = need
                slot[0] = gpu.Chromosome(t, chars, stream=stream)
            ch = slot[0]
            if sliced:
                for piece in lane_bam.iter_target(t, int(slice_reads), owned=True):
                    try:
                        if piece.n_reads:
                            with bus:
                                ch.push_reads(piece); ch.sync()
                    finally:
                        piece.free()
            else:
                with bus:
                    ch.push_reads(batch); ch.sync()
                batch.free()
            res = ch.finish()
            cnv = ch.cnv(params=prm)
            text[t] = hostlib.vcf_contig(prm, name, chars, res.snv, res.snv_ave_rd, res.ins, res.del_ev, res.sv_ev, cnv.calls)
            if ctx_out is not None:
                ctx_out[t] = hostlib.ctx_contig(prm, t, res.sv_ev)

        def lane():
            stream = gpu.stream_create() if n_lanes > 1 else None
            slot = [None, 0]
            try:
                with hostlib.Bam(bam_path) as lane_bam:                    # one reader per lane (the batcher seeks in its file)
                    while not errors:
                        with pick:
                            if not work:
                                return
                            t = work.pop(0)
                        one_contig(lane_bam, t, stream, slot)
            except BaseException as e:                                      # surfaced by the caller's thread below
                errors.append(e)
            finally:
                if slot[0] is not None:
                    slot[0].close(); inflight.release(slot[1])
                gpu.stream_destroy(stream)

        if n_lanes <= 1:
            lane()
        else:
            th = [threading.Thread(target=lane) for _ in range(n_lanes)]
            for x in th:
                x.start()
            for x in th:
                x.join()
        if errors:
            raise errors[0]
    return text, prm


def ctx_vcf_text(params: Params, target_names: List[str], per_contig: Dict[int, np.ndarray]) -> str:
    """Body of <out>.ctx.vcf: mate pairing over the translocation records of all contigs (all ranks), in contig order."""
    from .params import CTX_RECORD_DTYPE
    recs = [per_contig[t] for t in sorted(per_contig)]
    allrec = np.concatenate(recs) if recs else np.zeros(0, dtype=CTX_RECORD_DTYPE)
    return hostlib.ctx_vcf(params, target_names, allrec)


def write_vcf(path: str, per_contig: Dict[int, str], fasta_name: str = ""):
    """<out>: the reference's header block (GROM.c:20517-20565; `fasta_name` is what it prints as ##reference) + the contigs in BAM header order."""
    with open(path, "w") as f:
        f.write(hostlib.vcf_header(fasta_name, False))
        for t in sorted(per_contig):
            f.write(per_contig[t])


def write_ctx_vcf(path: str, body: str, fasta_name: str = ""):
    """<out>.ctx.vcf: the reference's header block (GROM.c:22639-22677) + the paired translocation records."""
    with open(path, "w") as f:
        f.write(hostlib.vcf_header(fasta_name, True))
        f.write(body)
