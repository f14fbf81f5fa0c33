"""The C host program (tools/grom_b200.c -> grom_b200/GROM_b200): GROM's command line and output files over the two libraries.
CPU part: option handling, header blocks and file naming against the live reference binary (skipped where oracle/_ref/GROM_ref is
absent).  GPU part: whole output files == the reference's (modulo ##fileDate / ##reference and the two uninitialised fields of the
reference's small-insertion records), single worker == two workers merged."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from util import GOLDEN, ROOT
from oracle import pyoracle as po
from tools import synth

EXE = os.path.join(ROOT, "grom_b200", "GROM_b200")


def _strip(lines):
    return [l for l in lines if not l.startswith("##fileDate=") and not l.startswith("##reference=")]


def _golden_fasta(tmp_path):
    fa = str(tmp_path / "g1.fa")
    with gzip.open(os.path.join(GOLDEN, "g1.fa.gz"), "rb") as f, open(fa, "wb") as o:
        shutil.copyfileobj(f, o)
    return fa


def _run(args, **kw):
    return subprocess.run([EXE] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)


def test_cli_contract_without_gpu(tmp_path):
    """required options, -h, unknown letters, file naming; --merge-only writes both files with the reference's header blocks"""
    assert os.path.exists(EXE), "build() compiles tools/grom_b200.c"
    assert _run(["-h"]).returncode == 0
    r = _run(["-r", "x.fa", "-o", "x.vcf"])
    assert r.returncode == 1 and "No bam file specified" in r.stderr
    r = _run(["-i", "x.bam", "-r", "x.fa"])
    assert r.returncode == 1 and "No output file specified" in r.stderr
    assert _run(["-i", "x.bam", "-r", "x.fa", "-o", "x.vcf", "-f"]).returncode == 1          # tab-separated debug output is not built
    fa = _golden_fasta(tmp_path)
    for out, ctx in (("o.vcf", "o.ctx.vcf"), ("o.txt", "o.txt.ctx")):                      # src/GROM.c:22431-22445
        r = _run(["-i", os.path.join(GOLDEN, "g1.bam"), "-r", fa, "-o", tmp_path / out, "--merge-only", "--libstats", "400,150,300,500"])
        assert r.returncode == 0, r.stderr
        assert os.path.exists(tmp_path / out) and os.path.exists(tmp_path / ctx)
    head = open(tmp_path / "o.vcf").read().splitlines()
    assert head[0] == "##fileformat=VCFv4.2" and head[2] == f"##reference={fa}" and head[-1] == "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT"
    assert sum(l.startswith("##FORMAT=") for l in head) == 35 and sum(l.startswith("##FORMAT=") for l in open(tmp_path / "o.ctx.vcf")) == 30


@pytest.mark.skipif(not po.have_reference("ref"), reason="oracle/_ref/GROM_ref not built")
def test_header_blocks_equal_the_live_reference(tmp_path):
    spec = synth.SynthSpec(contigs=[("chrA", 60_000), ("chrZ", 20_000)], depth=10, seed=3)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), synth.simulate(spec))
    po.run_reference(bam, fa, str(tmp_path / "ref.vcf"))
    r = _run(["-i", bam, "-r", fa, "-o", tmp_path / "mine.vcf", "--merge-only", "--libstats", "400,150,300,500"])
    assert r.returncode == 0, r.stderr
    for name in ("vcf", "ctx.vcf"):
        ref = _strip([l for l in open(tmp_path / f"ref.{name}") if l.startswith("#")])
        mine = _strip([l for l in open(tmp_path / f"mine.{name}") if l.startswith("#")])
        assert len(ref) > 30 and mine == ref


@pytest.mark.gpu
@pytest.mark.parametrize("tag,flags", [("default", []), ("rmdup", ["-M"])])
def test_c_host_reproduces_reference_files_golden(tmp_path, tag, flags):
    g = np.load(os.path.join(GOLDEN, f"g1_{tag}.npz"))
    fa = _golden_fasta(tmp_path)
    r = _run(["-i", os.path.join(GOLDEN, "g1.bam"), "-r", fa, "-o", tmp_path / "o.vcf", "--stats", tmp_path / "st.json"] + flags)
    assert r.returncode == 0, r.stderr + r.stdout
    mine = [l for l in open(tmp_path / "o.vcf") if not l.startswith("#")]
    ref = [l for l in str(g["vcf"]).splitlines(keepends=True) if not l.startswith("#")]
    assert len(ref) > 100 and po.normalise_records(mine) == po.normalise_records(ref)
    m = g["mean"]
    assert f"insert mean, insert minimum, insert maximum: {int(max(m[0], m[1]))} {int(m[2])} {int(m[3])}" in r.stdout       # the reference's own progress line
    # two workers (contigs split largest-first) + merge == one worker
    for rank in (0, 1):
        rr = _run(["-i", os.path.join(GOLDEN, "g1.bam"), "-r", fa, "-o", tmp_path / "p.vcf", "--rank", rank, "--world", 2, "--device", 0, "--parts-only"] + flags)
        assert rr.returncode == 0, rr.stderr
    parts = sorted(f for f in os.listdir(tmp_path) if f.startswith("p.vcf.part."))
    assert len(parts) == 3
    assert _run(["-i", os.path.join(GOLDEN, "g1.bam"), "-r", fa, "-o", tmp_path / "p.vcf", "--merge-only"] + flags).returncode == 0
    assert _strip(open(tmp_path / "p.vcf").readlines()) == _strip(open(tmp_path / "o.vcf").readlines())
    assert not [f for f in os.listdir(tmp_path) if ".part." in f or ".ctxpart." in f]


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_reference("ref"), reason="oracle/_ref/GROM_ref not built")
def test_c_host_whole_files_against_live_reference(tmp_path):
    """every record class incl. translocations, -M -q 30 -b 26: <out> and <out>.ctx.vcf of GROM_b200 against GROM_ref run on the same files"""
    spec = synth.SynthSpec(contigs=[("chrA", 400_000), ("chrB", 300_000), ("chrZ", 50_000)], depth=30, seed=71, dup_frac=0.05, sa_frac=0.5, disc_frac=0.02,
                           sv_sites_per_mb=12.0, munmap_frac=0.01, sv_classes=15, cnv_per_mb=2.0, cnv_min=15000, cnv_max=30000)
    fa, bam = synth.write_dataset(str(tmp_path / "d"), synth.simulate(spec))
    flags = ["-M", "-q", "30", "-b", "26"]
    po.run_reference(bam, fa, str(tmp_path / "ref.vcf"), args=flags, seed=1)
    r = _run(["-i", bam, "-r", fa, "-o", tmp_path / "mine.vcf"] + flags)
    assert r.returncode == 0, r.stderr + r.stdout
    ref = _strip(open(tmp_path / "ref.vcf").readlines()); mine = _strip(open(tmp_path / "mine.vcf").readlines())
    assert [l for l in mine if l.startswith("#")] == [l for l in ref if l.startswith("#")]
    rb, mb = [l for l in ref if not l.startswith("#")], [l for l in mine if not l.startswith("#")]
    assert len(rb) > 500 and po.normalise_records(mb) == po.normalise_records(rb)
    assert _strip(open(tmp_path / "mine.ctx.vcf").readlines()) == _strip(open(tmp_path / "ref.ctx.vcf").readlines())
    assert sum("SVTYPE=BND" in l for l in open(tmp_path / "mine.ctx.vcf")) >= 2
