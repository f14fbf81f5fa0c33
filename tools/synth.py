"""Synthetic FASTA + paired-end read generator (test / bench tooling, numpy only).

Produces, per contig, a `ReadBatch` in BAM order together with the reference
characters, so the same reads can be (a) pushed straight through the C-ABI and
(b) serialised with `hostlib.write_bam` for the reference binary.  The shapes
follow SURVEY.md §8(d): 2x150 bp reads, insert ~N(400,40), MAPQ 60 (97 %) / low
(3 %), base qualities from {37,37,37,30,25,12}, 0.2 % substitution errors,
planted SNVs, small indels, soft/hard clips, duplicates, discordant pairs of
every orientation class, mate-unmapped and inter-contig pairs, SA tags.

Nothing here is on the product path.
"""
from __future__ import annotations

import os
import re
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grom_b200.reads import (BASE_ALIGN, CDEL, CHARD_CLIP, CINS, CMATCH, CREF_SKIP, CSOFT_CLIP, CEQUAL, CDIFF,  # noqa: E402
                             FDUP, FMREVERSE, FMUNMAP, FPAIRED, FPROPER, FREAD1, FREAD2, FREVERSE, FUNMAP,
                             ReadBatch, fnv1a64)

_CODE_OF = np.zeros(256, dtype=np.uint8) + 15
for ch, code in (("A", 1), ("C", 2), ("G", 4), ("T", 8), ("a", 1), ("c", 2), ("g", 4), ("t", 8)):
    _CODE_OF[ord(ch)] = code
_ACGT_CODES = np.array([1, 2, 4, 8], dtype=np.uint8)


@dataclass
class SynthSpec:
    contigs: List[Tuple[str, int]]
    depth: float = 30.0
    read_len: int = 150
    ins_mean: float = 400.0
    ins_sd: float = 40.0
    ins_floor: int = 170            # < 2*read_len so that some mates overlap (name de-dup path)
    seed: int = 1
    snv_every: int = 1000
    indel_every: int = 10000
    err_rate: float = 0.002
    low_mapq_frac: float = 0.03
    clip_frac: float = 0.02
    hardclip_frac: float = 0.002
    dup_frac: float = 0.0           # exact PCR duplicates (pairs)
    flagdup_frac: float = 0.003     # reads carrying BAM_FDUP
    disc_frac: float = 0.01         # discordant pairs (all classes together)
    unpaired_frac: float = 0.004
    munmap_frac: float = 0.003
    refskip_frac: float = 0.0005
    sv_sites_per_mb: float = 1.0    # planted clustered deletions (pairs + split reads)
    n_frac: float = 0.01
    lower_frac: float = 0.1
    long_name_frac: float = 0.0005
    sa_frac: float = 0.5            # fraction of (>= 20 bp) soft-clipped reads that carry an SA tag
    simple: bool = False            # bench mode: only the vectorised read classes
    simple_disc_frac: float = 0.0   # simple mode: vectorised deletion-like / inverted / mate-unmapped pairs (bench realism)
    names: bool = True              # False: no read-name strings (hash only; such a batch cannot be written as BAM)
    cnv_per_mb: float = 0.0         # planted copy-number segments (alternating loss / gain of one copy, every 4th a full loss)
    cnv_min: int = 20_000
    cnv_max: int = 120_000
    sv_classes: float = 0.0         # planted clusters per Mb of EACH further class: tandem duplication (RF pairs), inversion (FF and RR pairs),
                                    # translocation (mates on another contig), insertion (soft clips + unmapped mates + short pairs)
    at_repeats: int = 0             # planted (AT)n runs of 24-60 bp whose coverage is thinned (exercises the biased-repeat path)
    at_repeat_keep: float = 0.3


def make_reference(length: int, rng: np.random.Generator, n_frac=0.01, lower_frac=0.1) -> np.ndarray:
    """ASCII reference with GC-heterogeneous segments, N runs >= 100 and soft-masked stretches."""
    seg = 50_000
    nseg = (length + seg - 1) // seg
    gc = rng.uniform(0.30, 0.65, nseg)
    u = rng.random(length, dtype=np.float32)
    g = np.repeat(gc, seg)[:length].astype(np.float32)
    is_gc = u < g
    pick = rng.integers(0, 2, length, dtype=np.uint8)
    chars = np.where(is_gc, np.where(pick == 0, ord("C"), ord("G")), np.where(pick == 0, ord("A"), ord("T"))).astype(np.uint8)
    # dinucleotide repeats (exercise the repeat pre-pass)
    for _ in range(max(1, length // 200_000)):
        a = int(rng.integers(0, max(1, length - 200)))
        n = int(rng.integers(24, 120))
        di = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 2, replace=False)
        chars[a:a + n] = np.tile(di, n // 2 + 1)[:n][: max(0, min(n, length - a))]
    n_total = int(length * n_frac)
    while n_total > 0 and length > 5000:
        run = int(rng.integers(100, 2000))
        a = int(rng.integers(1000, max(1001, length - run - 1000)))
        chars[a:a + run] = ord("N")
        n_total -= run
    low_total = int(length * lower_frac)
    while low_total > 0 and length > 5000:
        run = int(rng.integers(50, 3000))
        a = int(rng.integers(0, max(1, length - run)))
        chars[a:a + run] |= 0x20
        low_total -= run
    return chars


def write_fasta(path: str, contigs: List[Tuple[str, np.ndarray]], width: int = 60):
    with open(path, "wb") as f:
        for name, chars in contigs:
            f.write(b">" + name.encode() + b"\n")
            n = len(chars)
            full = n // width
            if full:
                body = np.empty((full, width + 1), dtype=np.uint8)
                body[:, :width] = chars[: full * width].reshape(full, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n % width:
                f.write(chars[full * width:].tobytes() + b"\n")


def _pack_nibbles(codes: np.ndarray) -> np.ndarray:
    """[n, S] 4-bit codes (S even) -> [n, S/2] BAM nibble bytes."""
    return ((codes[:, 0::2] << 4) | codes[:, 1::2]).astype(np.uint8)


@dataclass
class SynthContig:
    name: str
    chars: np.ndarray
    batch: ReadBatch
    truth: Dict[str, np.ndarray] = field(default_factory=dict)


def _simulate_contig(tid: int, name: str, length: int, spec: SynthSpec, n_contigs: int,
                     contig_lens: List[int], rng: np.random.Generator, chars: np.ndarray) -> SynthContig:
    rl = spec.read_len
    S = (rl + BASE_ALIGN - 1) // BASE_ALIGN * BASE_ALIGN
    up = chars & 0xDF
    refcode = _CODE_OF[up]
    isn = refcode == 15
    # two haplotypes with planted SNVs
    hap = [refcode.copy(), refcode.copy()]
    n_snv = max(1, length // spec.snv_every) if spec.snv_every else 0
    snv_pos = np.unique(rng.integers(2000, max(2001, length - 2000), n_snv)) if n_snv else np.zeros(0, dtype=np.int64)
    snv_pos = snv_pos[~isn[snv_pos]] if n_snv else snv_pos
    shift = rng.integers(1, 4, len(snv_pos))
    ref_idx = np.log2(np.maximum(refcode[snv_pos], 1)).astype(np.int64)
    alt = _ACGT_CODES[(ref_idx + shift) % 4]
    het = (np.arange(len(snv_pos)) % 2) == 1
    hap[0][snv_pos] = alt
    hap[1][snv_pos[~het]] = alt[~het]
    # N in the sample -> random base
    for h in hap:
        m = h == 15
        h[m] = _ACGT_CODES[rng.integers(0, 4, int(m.sum()))]

    n_pairs = int(spec.depth * length / (2 * rl))
    ins = np.clip(np.rint(rng.normal(spec.ins_mean, spec.ins_sd, n_pairs)), spec.ins_floor, None).astype(np.int64)
    fs = np.sort(rng.integers(0, max(1, length - int(ins.max()) - 1), n_pairs)).astype(np.int64)
    hp = rng.integers(0, 2, n_pairs)
    cnv_truth = []
    if spec.at_repeats and name in _AT_RUNS:
        # pairs whose first read overlaps a planted (AT)n run are mostly lost
        ra, rn = _AT_RUNS.pop(name)
        j = np.searchsorted(ra, fs + rl, side="left") - 1
        hit = (j >= 0) & (ra[np.maximum(j, 0)] + rn[np.maximum(j, 0)] > fs)
        keep = ~hit | (rng.random(n_pairs) < spec.at_repeat_keep)
        fs, ins, hp = fs[keep], ins[keep], hp[keep]
        n_pairs = len(fs)
    if spec.cnv_per_mb > 0:
        # copy-number segments: thin (loss) or thicken (gain) the pairs that start inside the segment
        keep = np.ones(n_pairs, dtype=bool)
        extra = []
        for k in range(max(1, int(length / 1e6 * spec.cnv_per_mb))):
            L = int(rng.integers(spec.cnv_min, spec.cnv_max))
            a = int(rng.integers(15_000, max(15_001, length - L - 15_000)))
            i0, i1 = np.searchsorted(fs, [a, a + L])
            if k % 2 == 0:
                keep[i0:i1] &= rng.random(i1 - i0) < (0.0 if k % 4 == 2 else 0.5)
                cnv_truth.append((a, a + L, 0 if k % 4 == 2 else 1))
            else:
                extra.append(rng.integers(a, a + L, (i1 - i0) // 2))
                cnv_truth.append((a, a + L, 3))
        fs = np.sort(np.concatenate([fs[keep]] + extra)).astype(np.int64)
        fs = np.minimum(fs, max(1, length - int(ins.max()) - 2))
        n_pairs = len(fs)
        ins = np.clip(np.rint(rng.normal(spec.ins_mean, spec.ins_sd, n_pairs)), spec.ins_floor, None).astype(np.int64)
        fs = np.minimum(fs, length - ins - 1)
        fs.sort()
        hp = rng.integers(0, 2, n_pairs)
    # PCR duplicates: copy (fs, ins, hap) of a random earlier pair
    if spec.dup_frac > 0:
        nd = int(n_pairs * spec.dup_frac)
        src = rng.integers(0, n_pairs, nd); dst = rng.choice(n_pairs, nd, replace=False)
        fs[dst] = fs[src]; ins[dst] = ins[src]; hp[dst] = hp[src]
        o = np.argsort(fs, kind="stable"); fs, ins, hp = fs[o], ins[o], hp[o]

    n = 2 * n_pairs
    pair_id = np.repeat(np.arange(n_pairs), 2)
    is2 = np.tile(np.array([0, 1]), n_pairs).astype(bool)
    pos = np.where(is2, np.repeat(fs + ins - rl, 2), np.repeat(fs, 2)).astype(np.int64)
    mpos = np.where(is2, np.repeat(fs, 2), np.repeat(fs + ins - rl, 2)).astype(np.int64)
    tlen = np.where(is2, -np.repeat(ins, 2), np.repeat(ins, 2)).astype(np.int64)
    flag = np.where(is2, FPAIRED | FPROPER | FREVERSE | FREAD2, FPAIRED | FPROPER | FMREVERSE | FREAD1).astype(np.int64)
    mtid = np.full(n, tid, dtype=np.int64)
    hpr = np.repeat(hp, 2)
    mapq = np.full(n, 60, dtype=np.int64)
    low = rng.random(n) < spec.low_mapq_frac
    mapq[low] = rng.integers(0, 20, int(low.sum()))
    l_qseq = np.full(n, rl, dtype=np.int64)

    # read bases straight from the haplotypes
    win0 = np.lib.stride_tricks.sliding_window_view(hap[0], rl)
    win1 = np.lib.stride_tricks.sliding_window_view(hap[1], rl)
    codes = np.zeros((n, S), dtype=np.uint8)
    m0 = hpr == 0
    codes[m0, :rl] = win0[pos[m0]]
    codes[~m0, :rl] = win1[pos[~m0]]
    # substitution errors
    n_err = rng.binomial(n * rl, spec.err_rate)
    er = rng.integers(0, n, n_err); ec = rng.integers(0, rl, n_err)
    codes[er, ec] = _ACGT_CODES[rng.integers(0, 4, n_err)]
    n_nb = max(1, n // 2000) if not spec.simple else 0
    codes[rng.integers(0, n, n_nb), rng.integers(0, rl, n_nb)] = 15          # a few N base calls
    qtab = np.array([37, 37, 37, 30, 25, 12], dtype=np.uint8)
    qual = np.zeros((n, S), dtype=np.uint8)
    qual[:, :rl] = qtab[rng.integers(0, 6, (n, rl), dtype=np.uint8)]

    cig: Dict[int, List[Tuple[int, int]]] = {}      # overrides (default: rl M)
    aux: Dict[int, bytes] = {}
    sa_info: Dict[int, tuple] = {}

    def mate(i):
        return i ^ 1

    if not spec.simple:
        # ---- discordant classes on whole pairs
        nd = int(n_pairs * spec.disc_frac)
        dp = rng.choice(n_pairs, nd, replace=False)
        cls = rng.integers(0, 6, nd)
        for p, c in zip(dp, cls):
            a, b = 2 * p, 2 * p + 1
            if c == 0:      # deletion-like: mate far downstream
                d = int(rng.integers(600, 20000))
                if pos[b] + d + rl < length:
                    pos[b] += d; mpos[a] = pos[b]; tlen[a] += d; tlen[b] -= d
                    flag[a] &= ~FPROPER; flag[b] &= ~FPROPER
                    codes[b, :rl] = hap[hpr[b]][pos[b]:pos[b] + rl]
            elif c == 1:    # FF
                flag[b] &= ~FREVERSE; flag[a] &= ~FMREVERSE; flag[a] &= ~FPROPER; flag[b] &= ~FPROPER
                d = int(rng.integers(0, 5000))
                if pos[b] + d + rl < length:
                    pos[b] += d; mpos[a] = pos[b]; tlen[a] += d; tlen[b] -= d
                    codes[b, :rl] = hap[hpr[b]][pos[b]:pos[b] + rl]
            elif c == 2:    # RR
                flag[a] |= FREVERSE; flag[b] |= FMREVERSE; flag[a] &= ~FPROPER; flag[b] &= ~FPROPER
                d = int(rng.integers(0, 5000))
                if pos[b] + d + rl < length:
                    pos[b] += d; mpos[a] = pos[b]; tlen[a] += d; tlen[b] -= d
                    codes[b, :rl] = hap[hpr[b]][pos[b]:pos[b] + rl]
            elif c == 3:    # RF (tandem-dup like)
                flag[a] |= FREVERSE; flag[a] &= ~FMREVERSE; flag[b] &= ~FREVERSE; flag[b] |= FMREVERSE
                flag[a] &= ~FPROPER; flag[b] &= ~FPROPER
                d = int(rng.integers(0, 5000))
                if pos[b] + d + rl < length:
                    pos[b] += d; mpos[a] = pos[b]; tlen[a] += d; tlen[b] -= d
                    codes[b, :rl] = hap[hpr[b]][pos[b]:pos[b] + rl]
            elif c == 4 and n_contigs > 1:    # mate on another contig: keep read a, drop b from this contig
                other = int((tid + 1 + rng.integers(0, n_contigs - 1)) % n_contigs)
                mtid[a] = other; mpos[a] = int(rng.integers(0, max(1, contig_lens[other] - rl))); tlen[a] = 0
                flag[a] &= ~FPROPER
                if rng.random() < 0.5:
                    flag[a] |= FREVERSE
                if rng.random() < 0.5:
                    flag[a] &= ~FMREVERSE
                flag[b] |= FUNMAP   # placeholder record, skipped by every consumer
            else:           # short insert
                d = int(rng.integers(0, 60))
                if pos[a] + d + rl <= pos[b] + rl:
                    pass
                tl = int(rng.integers(rl, 260))
                pos[b] = pos[a] + tl - rl; mpos[a] = pos[b]; tlen[a] = tl; tlen[b] = -tl
                codes[b, :rl] = hap[hpr[b]][pos[b]:pos[b] + rl]
        # ---- mate unmapped
        mu = rng.choice(n_pairs, int(n_pairs * spec.munmap_frac), replace=False)
        for p in mu:
            a, b = 2 * p, 2 * p + 1
            k, o = (a, b) if rng.random() < 0.5 else (b, a)
            flag[k] |= FMUNMAP; flag[k] &= ~FPROPER; tlen[k] = 0; mpos[k] = pos[k]
            flag[o] |= FUNMAP; flag[o] &= ~FPROPER; pos[o] = pos[k]; mpos[o] = pos[k]; tlen[o] = 0; mapq[o] = 0
        # ---- unpaired
        upi = rng.choice(n, int(n * spec.unpaired_frac), replace=False)
        flag[upi] &= FREVERSE
        # ---- BAM_FDUP
        fd = rng.choice(n, int(n * spec.flagdup_frac), replace=False)
        flag[fd] |= FDUP
        # ---- planted small indels carried by hap 0 (het) or both (hom)
        if spec.indel_every:
            n_ind = max(1, length // spec.indel_every)
            ipos = np.unique(rng.integers(3000, max(3001, length - 3000), n_ind))
            ilen = rng.integers(1, 21, len(ipos)); idel = rng.random(len(ipos)) < 0.5
            ihom = rng.random(len(ipos)) < 0.4
            iseq = [_ACGT_CODES[rng.integers(0, 4, int(k))] for k in ilen]
            order = np.argsort(pos, kind="stable")
            spos = pos[order]
            for j, ip in enumerate(ipos):
                lo = np.searchsorted(spos, ip - rl + 5, side="left"); hi = np.searchsorted(spos, ip - 5, side="right")
                for i in order[lo:hi]:
                    if i in cig or (flag[i] & FUNMAP) or not (ihom[j] or hpr[i] == 0):
                        continue
                    left = int(ip - pos[i])           # aligned bases before the event
                    if left < 5 or left > rl - 5:
                        continue
                    h = hap[hpr[i]]
                    if idel[j]:
                        right = rl - left
                        if ip + ilen[j] + right >= length:
                            continue
                        codes[i, :left] = h[pos[i]:pos[i] + left]
                        codes[i, left:rl] = h[ip + ilen[j]: ip + ilen[j] + right]
                        cig[i] = [(CMATCH, left), (CDEL, int(ilen[j])), (CMATCH, right)]
                    else:
                        k = int(min(ilen[j], rl - left - 3))
                        if k < 1:
                            continue
                        right = rl - left - k
                        codes[i, :left] = h[pos[i]:pos[i] + left]
                        codes[i, left:left + k] = iseq[j][:k]
                        codes[i, left + k:rl] = h[ip: ip + right]
                        cig[i] = [(CMATCH, left), (CINS, k), (CMATCH, right)]
        # ---- soft clips (with SA tags on half of them), hard clips, ref skips
        ci = rng.choice(n, int(n * spec.clip_frac), replace=False)
        for i in ci:
            if i in cig or (flag[i] & FUNMAP):
                continue
            k = int(rng.integers(1, 60)); leftside = rng.random() < 0.5
            junk = _ACGT_CODES[rng.integers(0, 4, k)]
            if leftside:
                codes[i, :k] = junk; cig[i] = [(CSOFT_CLIP, k), (CMATCH, rl - k)]
                pos[i] += k; mpos[mate(i)] = pos[i] if not (flag[mate(i)] & FUNMAP) and mtid[mate(i)] == tid and not (flag[mate(i)] & FMUNMAP) else mpos[mate(i)]
            else:
                codes[i, rl - k:rl] = junk; cig[i] = [(CMATCH, rl - k), (CSOFT_CLIP, k)]
            if rng.random() < spec.sa_frac and k >= 20:
                sa_pos = int(pos[i] + rng.integers(-3000, 3000)); sa_pos = max(1, sa_pos)
                strand = "+" if ((flag[i] & FREVERSE) == 0) == (rng.random() < 0.8) else "-"
                sa_cig = f"{rl - k}S{k}M" if not leftside else f"{k}M{rl - k}S"
                if rng.random() < 0.2:
                    sa_cig = f"{rl - k}S{k - 2}M2D2M" if not leftside else f"{k - 3}M1I2M{rl - k}S"
                sa_name = name if rng.random() < 0.85 else ("q" + name)
                sa_mq = int(rng.integers(0, 61))
                tag = f"{sa_name},{sa_pos},{strand},{sa_cig},{sa_mq},0;"
                pre = b""
                if rng.random() < 0.1:      # > 99 aux bytes: the reference then ignores the tag (src/GROM.c:5763)
                    pre = b"XXZ" + b"y" * 90 + b"\0"
                aux[i] = pre + b"SAZ" + tag.encode() + b"\0"
                if not pre:
                    ops = re.findall(r"(\d+)([A-Z])", sa_cig)
                    sa_info[i] = (sa_pos, 0 if strand == "+" else 1, sa_mq, 1 if sa_name.startswith(name) else 0,
                                  int(ops[0][0]) if ops[0][1] == "S" else 0, int(ops[-1][0]) if ops[-1][1] == "S" else 0,
                                  sum(int(a) for a, b in ops if b == "I") - sum(int(a) for a, b in ops if b == "D"))
        hi_ = rng.choice(n, int(n * spec.hardclip_frac), replace=False)
        for i in hi_:
            if i in cig or (flag[i] & FUNMAP):
                continue
            k = int(rng.integers(5, 50))
            if rng.random() < 0.5:
                codes[i, :rl - k] = codes[i, k:rl].copy(); qual[i, :rl - k] = qual[i, k:rl].copy()
                pos[i] += k; cig[i] = [(CHARD_CLIP, k), (CMATCH, rl - k)]
                if not (flag[mate(i)] & (FUNMAP | FMUNMAP)) and mtid[mate(i)] == tid:
                    mpos[mate(i)] = pos[i]
            else:
                cig[i] = [(CMATCH, rl - k), (CHARD_CLIP, k)]
            codes[i, rl - k:] = 0; qual[i, rl - k:] = 0
            l_qseq[i] = rl - k
        ri = rng.choice(n, int(n * spec.refskip_frac), replace=False)
        for i in ri:
            if i in cig or (flag[i] & FUNMAP):
                continue
            left = int(rng.integers(20, rl - 20)); skip = int(rng.integers(50, 400))
            if pos[i] + rl + skip >= length:
                continue
            h = hap[hpr[i]]
            codes[i, left:rl] = h[pos[i] + left + skip: pos[i] + skip + rl]
            cig[i] = [(CEQUAL, 10), (CDIFF, 1), (CMATCH, left - 11), (CREF_SKIP, skip), (CMATCH, rl - left)]
        # ---- planted clustered deletions: several supporting pairs + split reads
        n_sv = int(length / 1e6 * spec.sv_sites_per_mb)
        truth_sv = []
        for _ in range(n_sv):
            a = int(rng.integers(20000, max(20001, length - 60000))); dl = int(rng.integers(800, 30000))
            truth_sv.append((a, a + dl))
            lo = np.searchsorted(fs, a - 380); hi2 = np.searchsorted(fs, a - 160)
            for p in range(lo, hi2):
                if hp[p] != 0:
                    continue
                ia, ib = 2 * p, 2 * p + 1
                if ia in cig or ib in cig or (flag[ia] | flag[ib]) & (FUNMAP | FMUNMAP) or mtid[ia] != tid:
                    continue
                if pos[ia] + rl <= a and pos[ib] >= a - 40:
                    npos = pos[ib] + dl
                    if npos + rl >= length:
                        continue
                    pos[ib] = npos; mpos[ia] = npos; tlen[ia] = npos + rl - pos[ia]; tlen[ib] = -tlen[ia]
                    flag[ia] = FPAIRED | FMREVERSE | FREAD1; flag[ib] = FPAIRED | FREVERSE | FREAD2
                    codes[ib, :rl] = hap[0][npos:npos + rl]
        if spec.sv_classes > 0:
            other = (tid + 1) % n_contigs
            n_each = max(1, int(length / 1e6 * spec.sv_classes))
            for k in range(n_each * 5):
                kind = k % 5
                a = int(rng.integers(20000, max(20001, length - 60000))); dl = int(rng.integers(1500, 20000))
                tgt = int(rng.integers(5000, max(5001, contig_lens[other] - 5000)))
                lo = np.searchsorted(fs, a - 380); hi2 = np.searchsorted(fs, a - (20 if kind == 4 else 100))
                for p in range(lo, hi2):
                    ia, ib = 2 * p, 2 * p + 1
                    if ia in cig or ib in cig or (flag[ia] | flag[ib]) & (FUNMAP | FMUNMAP) or mtid[ia] != tid or not (flag[ia] & FPROPER):
                        continue
                    if kind == 4:
                        # insertion breakpoint at a: forward reads start there behind a clipped head, reverse reads end there before a
                        # clipped tail, both without a mapped mate (the mate sits in the inserted sequence)
                        if pos[ia] < a < pos[ia] + rl - 20 and a - pos[ia] >= 20:
                            k0 = a - pos[ia]
                            cig[ia] = [(CSOFT_CLIP, k0), (CMATCH, rl - k0)]
                            codes[ia, k0:rl] = hap[0][a:a + rl - k0]
                            pos[ia] = a
                            flag[ia] = FPAIRED | FMUNMAP | FREAD1; tlen[ia] = 0; mpos[ia] = a
                            flag[ib] = FPAIRED | FUNMAP | FREAD2; pos[ib] = a; mpos[ib] = a; tlen[ib] = 0
                        elif pos[ib] < a < pos[ib] + rl - 20 and a - pos[ib] >= 20:
                            cig[ib] = [(CMATCH, a - pos[ib]), (CSOFT_CLIP, rl - (a - pos[ib]))]
                            flag[ib] = FPAIRED | FREVERSE | FMUNMAP | FREAD2; tlen[ib] = 0; mpos[ib] = pos[ib]
                            flag[ia] = FPAIRED | FUNMAP | FREAD1; pos[ia] = pos[ib]; mpos[ia] = pos[ib]; tlen[ia] = 0
                        continue
                    if not (pos[ia] + rl <= a and pos[ib] >= a - 60):
                        continue
                    if kind == 0:        # tandem duplication of [a - dl, a): the reverse mate lands near the start of the copy (RF)
                        npos = a - dl + int(pos[ib] - (a - 60))
                        if npos < 1000:
                            continue
                        pos[ib] = npos; mpos[ia] = npos; tlen[ib] = pos[ia] + rl - npos; tlen[ia] = -tlen[ib]
                        flag[ia] = FPAIRED | FMREVERSE | FREAD1; flag[ib] = FPAIRED | FREVERSE | FREAD2
                    elif kind == 1:      # inversion, left breakpoint: both reads forward
                        npos = a + dl - int(pos[ib] - (a - 60)) - rl
                        if npos + rl >= length:
                            continue
                        pos[ib] = npos; mpos[ia] = npos; tlen[ia] = npos + rl - pos[ia]; tlen[ib] = -tlen[ia]
                        flag[ia] = FPAIRED | FREAD1; flag[ib] = FPAIRED | FREAD2
                    elif kind == 2:      # inversion, right breakpoint: both reads reverse
                        npos = a + dl + int(pos[ib] - (a - 60))
                        if npos + rl >= length:
                            continue
                        pos[ia], pos[ib] = pos[ib], npos
                        mpos[ia] = pos[ib]; mpos[ib] = pos[ia]; tlen[ia] = pos[ib] + rl - pos[ia]; tlen[ib] = -tlen[ia]
                        flag[ia] = FPAIRED | FREVERSE | FMREVERSE | FREAD1; flag[ib] = FPAIRED | FREVERSE | FMREVERSE | FREAD2
                        codes[ia, :rl] = hap[0][pos[ia]:pos[ia] + rl]
                    else:                # translocation: both mates are reported on the other contig
                        j = int(rng.integers(0, 120))
                        if other > tid and (other, a, tgt, tid) not in _CTX_PLANTS.setdefault(spec.seed, []):
                            _CTX_PLANTS[spec.seed].append((other, a, tgt, tid))      # the other contig gets the reciprocal reads
                        mtid[ia] = other; mpos[ia] = tgt + j; tlen[ia] = 0; flag[ia] = FPAIRED | FMREVERSE | FREAD1
                        mtid[ib] = other; mpos[ib] = tgt + 5000 + j; tlen[ib] = 0; flag[ib] = FPAIRED | FREVERSE | FREAD2
                    codes[ib, :rl] = hap[0][pos[ib]:pos[ib] + rl]
            # reciprocal side of the translocations planted on earlier contigs: reverse reads just after the target point whose mates
            # are reported back on the source contig, just before its breakpoint
            for (dst, a_src, tgt_here, src) in [x for x in _CTX_PLANTS.get(spec.seed, []) if len(x) == 4 and x[0] == tid]:
                lo = np.searchsorted(fs, tgt_here - 300); hi2 = np.searchsorted(fs, tgt_here - 100)
                for p in range(lo, hi2):
                    ia, ib = 2 * p, 2 * p + 1
                    if ia in cig or ib in cig or (flag[ia] | flag[ib]) & (FUNMAP | FMUNMAP) or mtid[ia] != tid or not (flag[ia] & FPROPER):
                        continue
                    if not (tgt_here <= pos[ib] < tgt_here + 130):
                        continue
                    mtid[ib] = src; mpos[ib] = a_src - 300 + int(pos[ib] - tgt_here); tlen[ib] = 0; flag[ib] = FPAIRED | FREVERSE | FREAD2
                    mtid[ia] = src; mpos[ia] = a_src + 5000; tlen[ia] = 0; flag[ia] = FPAIRED | FMREVERSE | FREAD1
        truth = {"snv_pos": snv_pos, "snv_het": het, "sv": np.array(truth_sv, dtype=np.int64).reshape(-1, 2), "cnv": cnv_truth}
    else:
        truth = {"snv_pos": snv_pos, "snv_het": het, "cnv": cnv_truth}
        if spec.simple_disc_frac > 0:
            # vectorised discordant classes: 60 % deletion-like (mate 2-20 kb downstream), 20 % same-strand (FF), 20 % mate unmapped
            nd = int(n_pairs * spec.simple_disc_frac)
            dp = rng.choice(n_pairs, nd, replace=False)
            kind = rng.random(nd)
            a_i, b_i = 2 * dp, 2 * dp + 1
            far = kind < 0.8
            d = rng.integers(2000, 20000, nd)
            ok = far & (pos[b_i] + d + rl < length)
            pos[b_i[ok]] += d[ok]; mpos[a_i[ok]] = pos[b_i[ok]]; tlen[a_i[ok]] += d[ok]; tlen[b_i[ok]] -= d[ok]
            flag[a_i[ok]] &= ~FPROPER; flag[b_i[ok]] &= ~FPROPER
            ff = ok & (kind >= 0.6)
            flag[b_i[ff]] &= ~FREVERSE; flag[a_i[ff]] &= ~FMREVERSE
            mb = b_i[ok]
            codes[mb, :rl] = np.where((hpr[mb] == 0)[:, None], win0[pos[mb]], win1[pos[mb]])
            mu = ~far
            flag[a_i[mu]] |= FMUNMAP; flag[a_i[mu]] &= ~FPROPER; tlen[a_i[mu]] = 0; mpos[a_i[mu]] = pos[a_i[mu]]
            flag[b_i[mu]] |= FUNMAP; flag[b_i[mu]] &= ~FPROPER; pos[b_i[mu]] = pos[a_i[mu]]; mpos[b_i[mu]] = pos[a_i[mu]]; tlen[b_i[mu]] = 0

    # names
    order = np.argsort(pos, kind="stable")
    n_cigar = np.ones(n, dtype=np.int64)
    for i, c in cig.items():
        n_cigar[i] = len(c)
    # sort everything
    pos, mpos, tlen, flag, mtid, mapq, l_qseq, pair_id, n_cigar = (x[order] for x in (pos, mpos, tlen, flag, mtid, mapq, l_qseq, pair_id, n_cigar))
    codes = codes[order]; qual = qual[order]
    inv = np.empty(n, dtype=np.int64); inv[order] = np.arange(n)
    cigar_off = np.concatenate([[0], np.cumsum(n_cigar)])[:-1]
    cigar = np.zeros(int(n_cigar.sum()), dtype=np.uint32)
    cigar[cigar_off] = (l_qseq.astype(np.uint32) << 4) | CMATCH
    for i, c in cig.items():
        j = inv[i]; o = int(cigar_off[j])
        for k, (op, ln) in enumerate(c):
            cigar[o + k] = (ln << 4) | op
    # read names "<contig>.<pair>" (+ a few >= 50 chars long)
    if not spec.names:
        # bench mode: no name strings, only a 64-bit per-pair id hash (both mates share it, like a shared read name)
        with np.errstate(over="ignore"):
            z = (pair_id.astype(np.uint64) + np.uint64(tid + 1) * np.uint64(0x9E3779B97F4A7C15))
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        batch = ReadBatch(
            tid=tid, pos=pos, mpos=mpos, tlen=tlen, mtid=mtid, l_qseq=l_qseq, flag=flag, n_cigar=n_cigar, mapq=mapq,
            qname_len=np.full(n, 12, dtype=np.uint8), qname_hash=z, cigar_off=cigar_off.astype(np.uint64),
            base_off=(np.arange(n, dtype=np.uint64) * np.uint64(S)), cigar=cigar, seq4=_pack_nibbles(codes).reshape(-1),
            qual=qual.reshape(-1), **_sa_arrays(n, sa_info, inv)).normalise()
        return SynthContig(name=name, chars=chars, batch=batch, truth=truth)
    base = np.char.add(f"{name}.", pair_id.astype(str))
    if not spec.simple and spec.long_name_frac > 0:
        lp = rng.choice(n_pairs, max(1, int(n_pairs * spec.long_name_frac)), replace=False)
        long_mask = np.isin(pair_id, lp)
        base = np.where(long_mask, np.char.add(base, "_" + "L" * 52), base)
    enc = np.char.encode(base, "ascii")
    lens = np.char.str_len(enc).astype(np.int64) + 1
    qname_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    pool = np.frombuffer(b"\0".join(enc.tolist()) + b"\0", dtype=np.uint8).copy()
    qname_hash = fnv1a64(pool, qname_off)
    # aux
    aux_off = np.zeros(n + 1, dtype=np.uint64)
    aux_bytes = bytearray()
    sa = _sa_arrays(n, sa_info, inv)
    if aux:
        alen = np.zeros(n, dtype=np.int64)
        items = sorted(((int(inv[i]), v) for i, v in aux.items()))
        for j, v in items:
            alen[j] = len(v)
        aux_off = np.concatenate([[0], np.cumsum(alen)]).astype(np.uint64)
        for j, v in items:
            aux_bytes += v
    batch = ReadBatch(
        tid=tid, pos=pos, mpos=mpos, tlen=tlen, mtid=mtid, l_qseq=l_qseq, flag=flag, n_cigar=n_cigar, mapq=mapq,
        qname_len=np.minimum(lens - 1, 255), qname_hash=qname_hash, cigar_off=cigar_off.astype(np.uint64),
        base_off=(np.arange(n, dtype=np.uint64) * np.uint64(S)), cigar=cigar, seq4=_pack_nibbles(codes).reshape(-1),
        qual=qual.reshape(-1), qname_off=qname_off, qname_pool=pool,
        aux_off=aux_off, aux_pool=np.frombuffer(bytes(aux_bytes), dtype=np.uint8).copy(), **sa).normalise()
    return SynthContig(name=name, chars=chars, batch=batch, truth=truth)


def _sa_arrays(n: int, sa_info: Dict[int, tuple], inv: np.ndarray) -> Dict[str, np.ndarray]:
    """the pre-parsed first SA entry per read (what the host batcher extracts from the aux block), in sorted read order"""
    sa = dict(sa_pos=np.full(n, -1, dtype=np.int32), sa_start_adj=np.zeros(n, np.int32), sa_end_adj=np.zeros(n, np.int32),
              sa_end_adj_indel=np.zeros(n, np.int32), sa_strand=np.zeros(n, np.uint8), sa_mapq=np.full(n, -1, np.int16),
              sa_same_chr=np.zeros(n, np.uint8))
    for i, v in sa_info.items():
        j = int(inv[i])
        sa["sa_pos"][j], sa["sa_strand"][j], sa["sa_mapq"][j], sa["sa_same_chr"][j] = v[0], v[1], v[2], v[3]
        sa["sa_start_adj"][j], sa["sa_end_adj"][j], sa["sa_end_adj_indel"][j] = v[4], v[5], v[6]
    return sa


_AT_RUNS: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}
_CTX_PLANTS: Dict[int, list] = {}


def simulate(spec: SynthSpec, references: Optional[Dict[str, np.ndarray]] = None) -> List[SynthContig]:
    """`references` = {contig name: uint8 characters}: reads are simulated on these sequences (case and N runs as they are) instead of
    a random one -- e.g. the reference's own tilapia FASTA."""
    rng = np.random.default_rng(spec.seed)
    _CTX_PLANTS.pop(spec.seed, None)
    lens = [l for _, l in spec.contigs]
    out = []
    for tid, (name, length) in enumerate(spec.contigs):
        if references is not None and name in references:
            chars = np.ascontiguousarray(references[name], dtype=np.uint8).copy()
            assert len(chars) == length, f"{name}: {len(chars)} characters given, {length} in the spec"
        else:
            chars = make_reference(length, rng, spec.n_frac, spec.lower_frac)
        if spec.at_repeats and length > 20_000:
            a = np.sort(rng.integers(5_000, length - 5_000, spec.at_repeats))
            a = a[np.concatenate([[True], np.diff(a) > 700])]
            n = rng.integers(24, 60, len(a))
            for x, k in zip(a, n):
                chars[x:x + k] = np.tile(np.frombuffer(b"AT", dtype=np.uint8), k // 2 + 1)[:k]
            _AT_RUNS[name] = (a, n)
        out.append(_simulate_contig(tid, name, length, spec, len(spec.contigs), lens, rng, chars))
    return out


def write_dataset(prefix: str, contigs: List[SynthContig], level: int = 1) -> Tuple[str, str]:
    """Write <prefix>.fa and <prefix>.bam(+.bai); returns (fasta, bam)."""
    from grom_b200 import hostlib
    fa, bam = prefix + ".fa", prefix + ".bam"
    write_fasta(fa, [(c.name, c.chars) for c in contigs])
    hostlib.write_bam(bam, [c.name for c in contigs], [len(c.chars) for c in contigs], [c.batch for c in contigs], level)
    for ext in (".mean", ".info"):
        for p in (bam + ext, fa + ext):
            if os.path.exists(p):
                os.remove(p)
    return fa, bam


def batch_from_records(tid: int, recs: List[dict]) -> ReadBatch:
    """Hand-made reads for edge-case tests.  Each record: pos, cigar [(op,len)...], seq (str over ACGTN=...),
    qual (int or list), and optional flag, mapq, mpos, mtid, tlen, name, sa=(pos,strand,mapq,same,start_adj,end_adj,indel)."""
    n = len(recs)
    recs = sorted(recs, key=lambda r: r["pos"])
    from grom_b200.reads import NT16
    lq = np.array([len(r["seq"]) for r in recs], dtype=np.int64)
    slots = (lq + BASE_ALIGN - 1) // BASE_ALIGN * BASE_ALIGN
    base_off = np.concatenate([[0], np.cumsum(slots)])[:-1] if n else np.zeros(0, dtype=np.int64)
    total = int(slots.sum())
    codes = np.zeros(total + (total & 1), dtype=np.uint8)
    qual = np.zeros(total + (total & 1), dtype=np.uint8)
    ncig = np.array([len(r["cigar"]) for r in recs], dtype=np.int64)
    cig_off = np.concatenate([[0], np.cumsum(ncig)])[:-1] if n else np.zeros(0, dtype=np.int64)
    cigar = np.zeros(int(ncig.sum()), dtype=np.uint32)
    names = []
    sa = dict(sa_pos=np.full(n, -1, dtype=np.int32), sa_start_adj=np.zeros(n, np.int32), sa_end_adj=np.zeros(n, np.int32),
              sa_end_adj_indel=np.zeros(n, np.int32), sa_strand=np.zeros(n, np.uint8), sa_mapq=np.full(n, -1, np.int16),
              sa_same_chr=np.zeros(n, np.uint8))
    for i, r in enumerate(recs):
        o = int(base_off[i])
        codes[o:o + lq[i]] = [NT16.index(ch) for ch in r["seq"]]
        q = r.get("qual", 30)
        qual[o:o + lq[i]] = q if not np.isscalar(q) else np.full(lq[i], q)
        for k, (op, ln) in enumerate(r["cigar"]):
            cigar[int(cig_off[i]) + k] = (ln << 4) | op
        names.append(r.get("name", f"r{i}"))
        if "sa" in r:
            s = r["sa"]
            sa["sa_pos"][i], sa["sa_strand"][i], sa["sa_mapq"][i], sa["sa_same_chr"][i] = s[0], s[1], s[2], s[3]
            sa["sa_start_adj"][i], sa["sa_end_adj"][i], sa["sa_end_adj_indel"][i] = s[4], s[5], s[6]
    enc = [s.encode() for s in names]
    lens = np.array([len(e) + 1 for e in enc], dtype=np.int64)
    qname_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    pool = np.frombuffer(b"\0".join(enc) + b"\0", dtype=np.uint8).copy() if n else np.zeros(0, dtype=np.uint8)
    g = lambda k, d: np.array([r.get(k, d) for r in recs], dtype=np.int64)  # noqa: E731
    return ReadBatch(
        tid=tid, pos=g("pos", 0), mpos=g("mpos", 0), tlen=g("tlen", 0), mtid=g("mtid", tid), l_qseq=lq, flag=g("flag", 0),
        n_cigar=ncig, mapq=g("mapq", 60), qname_len=np.minimum(lens - 1, 255), qname_hash=fnv1a64(pool, qname_off) if n else np.zeros(0, np.uint64),
        cigar_off=cig_off.astype(np.uint64), base_off=base_off.astype(np.uint64), cigar=cigar,
        seq4=((codes[0::2] << 4) | codes[1::2]).astype(np.uint8), qual=qual, qname_off=qname_off, qname_pool=pool,
        aux_off=np.zeros(n + 1, dtype=np.uint64), aux_pool=np.zeros(0, dtype=np.uint8), **sa).normalise()


def slice_batch(b: ReadBatch, i0: int, i1: int) -> ReadBatch:
    """Reads [i0, i1) of a batch as a self-contained batch (offsets rebased) -- for multi-push tests."""
    if i1 <= i0:
        raise ValueError("empty slice")
    c0 = int(b.cigar_off[i0]); c1 = int(b.cigar_off[i1 - 1]) + int(b.n_cigar[i1 - 1])
    s0 = int(b.base_off[i0])
    s1 = int(b.base_off[i1 - 1]) + (int(b.l_qseq[i1 - 1]) + BASE_ALIGN - 1) // BASE_ALIGN * BASE_ALIGN
    kw = {k: getattr(b, k)[i0:i1].copy() for k in ["pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len",
                                                   "qname_hash", "sa_pos", "sa_start_adj", "sa_end_adj", "sa_end_adj_indel",
                                                   "sa_strand", "sa_mapq", "sa_same_chr"]}
    return ReadBatch(tid=b.tid, cigar_off=b.cigar_off[i0:i1] - np.uint64(c0), base_off=b.base_off[i0:i1] - np.uint64(s0),
                     cigar=b.cigar[c0:c1].copy(), seq4=b.seq4[s0 // 2:s1 // 2].copy(), qual=b.qual[s0:s1].copy(), **kw).normalise()


def concat_batches(parts: List[ReadBatch]) -> ReadBatch:
    """Consecutive pieces of one target (each self-contained, offsets from 0) as one batch -- the inverse of slice_batch."""
    parts = [p for p in parts if p.n_reads]
    if not parts:
        raise ValueError("no reads")
    per_read = ["pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len", "qname_hash", "sa_pos", "sa_start_adj", "sa_end_adj",
                "sa_end_adj_indel", "sa_strand", "sa_mapq", "sa_same_chr"]
    kw = {k: np.concatenate([getattr(p, k) for p in parts]) for k in per_read}
    c_at = np.cumsum([0] + [len(p.cigar) for p in parts])
    s_at = np.cumsum([0] + [len(p.qual) for p in parts])
    return ReadBatch(tid=parts[0].tid,
                     cigar_off=np.concatenate([p.cigar_off + np.uint64(c_at[i]) for i, p in enumerate(parts)]),
                     base_off=np.concatenate([p.base_off + np.uint64(s_at[i]) for i, p in enumerate(parts)]),
                     cigar=np.concatenate([p.cigar for p in parts]), seq4=np.concatenate([p.seq4[:len(p.qual) // 2] for p in parts]),
                     qual=np.concatenate([p.qual for p in parts]), **kw).normalise()
