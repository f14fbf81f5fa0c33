"""Config 4 (whole genome, contigs assigned largest-first to the GPUs): the record text of every contig must not depend on the number
of ranks.  bench.py --workload wgs prints a digest over the per-contig VCF text in BAM header order; one rank against two ranks
(both on cuda:0 here, gloo for the gather -- the GPU test box has one GPU) must agree, and so must the LPT plan with partition.py."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT


def _run(n, port):
    args = ["bench.py", "--workload", "wgs", "--wgs-scale", "0.001", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--gpus", str(n)]
    if n == 1:
        cmd = [sys.executable] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    env = dict(os.environ, GROM_DIST_BACKEND="gloo")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.gpu
def test_wgs_digest_does_not_depend_on_the_number_of_ranks():
    one, two = _run(1, 0), _run(2, 29551)
    assert one["scaling"] == two["scaling"] == "strong" and one["n_gpus"] == 1 and two["n_gpus"] == 2
    assert one["records"] > 1000 and one["results"]["cnv_calls"] > 0
    assert one["genome_digest"] == two["genome_digest"] and one["records"] == two["records"] and one["results"] == two["results"]
    from grom_b200.partition import assign_contigs
    from tools.wgs_bench import genome
    g = genome(0.001)
    plan = assign_contigs([l for _, l in g], 2)
    assert two["lpt"]["contigs_per_rank"] == [[g[t][0] for t in p] for p in plan]
    assert sorted(sum(two["lpt"]["contigs_per_rank"], [])) == sorted(n for n, _ in g)
    assert two["lpt"]["imbalance"] < 1.1
