"""The drop-in boundary: the shared libraries load, export every symbol the headers declare, and the CUDA
library refuses to run (loudly) without a device -- there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

from util import ROOT, tables
from grom_b200 import gpu, hostlib
from grom_b200.params import Params


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", txt)))


def test_gromgpu_exports_every_declared_symbol():
    names = _declared("gromgpu.h", "gromgpu_")
    assert len(names) >= 13
    L = gpu.lib()
    for n in names:
        assert hasattr(L, n), f"libgromgpu.so does not export {n}"


def test_gromhost_exports_every_declared_symbol():
    names = _declared("gromhost.h", "gromhost_")
    L = hostlib.lib()
    for n in names:
        assert hasattr(L, n), f"libgromhost.so does not export {n}"


def test_integration_doc_lists_every_declared_symbol():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for header, prefix in (("gromgpu.h", "gromgpu_"), ("gromhost.h", "gromhost_")):
        for n in _declared(header, prefix):
            assert "`" + n + "`" in doc, f"INTEGRATION.md does not mention {n}"


def test_params_struct_matches_header():
    # sizeof(grom_params) as the C compiler sees it == ctypes mirror
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "grom_params.h"\n#include "grom_reads.h"\nint main(){printf("%zu %zu %zu %d", sizeof(grom_params), sizeof(grom_snv_cand), sizeof(grom_read_batch), GA_COUNT);return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    from grom_b200.params import GA_COUNT, SNV_CAND_DTYPE
    from grom_b200.reads import CReadBatch
    assert int(out[0]) == C.sizeof(Params)
    assert int(out[1]) == SNV_CAND_DTYPE.itemsize
    assert int(out[2]) == C.sizeof(CReadBatch)
    assert int(out[3]) == GA_COUNT


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    hez, mq = tables()
    with pytest.raises(gpu.GromGpuError, match="no CUDA device|no CPU fallback"):
        gpu.init(0, hez, mq, Params.default())
