"""N > 1 path on CPU: two gloo ranks partition the contigs of the golden BAM largest-first (the reference's -P
policy), each runs the ORACLE on its own contigs (stand-in for the GPU kernels, this is a host-logic test), the
per-contig SNV text is gathered and merged in BAM header order; the result must equal the single-rank output and no
data-path collective is used (only the final gather of text)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT, golden_batches, golden_params, load_golden_fasta, tables_7digit, GOLDEN
from grom_b200.partition import assign_contigs, merge_in_header_order, skip_contig


def test_assign_contigs_lpt():
    lens = [248, 242, 198, 190, 181, 170, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    for ws in (1, 2, 4, 8):
        ranks = assign_contigs(lens, ws)
        flat = sorted(i for r in ranks for i in r)
        assert flat == list(range(len(lens)))
        loads = [sum(lens[i] for i in r) for r in ranks]
        assert max(loads) - min(loads) <= max(lens)            # LPT bound
        for r in ranks:                                         # each rank processes largest first
            assert [lens[i] for i in r] == sorted((lens[i] for i in r), reverse=True)
    assert assign_contigs([5, 5, 5], 2) == [[0, 2], [1]]
    assert skip_contig("chrY", 0) and not skip_contig("chrY", 1) and not skip_contig("chr1", 0)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, batches = golden_batches()
    fasta = load_golden_fasta()
    hez, mq = tables_7digit()
    g = np.load(os.path.join(GOLDEN, "g1_default.npz"))
    prm = golden_params(g, 0)
    mine = assign_contigs([len(fasta[n]) for n in names], world)[rank]
    texts = {}
    for tid in mine:
        r = po.run_chr(prm, batches[tid], fasta[names[tid]], hez, mq)
        texts[tid] = po.format_snv_vcf(prm, names[tid].lower(), fasta[names[tid]], r.snv, r.snv_ave_rd)
    gathered = [None] * world
    dist.all_gather_object(gathered, texts)
    if rank == 0:
        open(os.path.join(out_dir, "merged.txt"), "w").write(merge_in_header_order(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_matches_single_rank(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    merged = open(tmp_path / "merged.txt").read().splitlines(keepends=True)
    g = np.load(os.path.join(GOLDEN, "g1_default.npz"))
    ref = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#") and l.split("\t")[2] == ""]
    assert len(merged) > 20 and merged == ref            # == the reference's own SNV records, in its contig order
