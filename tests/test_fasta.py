"""FASTA index + load of libgromhost (grom_b200/host/fasta.c) against a character-by-character restatement of the reference's two passes
(fgets into a 1000-byte buffer; src/GROM.c:1332-1417 and 21011-21045), on well-formed files and on the corner cases where the reference's
line rules show: CRLF, trailing blanks, a short last line, blank lines, lines longer than the fgets buffer, headers with descriptions,
long names, an empty header, no final newline."""
import os

import numpy as np

from grom_b200 import hostlib
from grom_b200.pipeline import read_fasta
from tools import synth


def fgets_lines(data: bytes):
    p = 0
    while p < len(data):
        q = data.find(b"\n", p, p + 999)
        e = q + 1 if q >= 0 else min(len(data), p + 999)
        yield p, data[p:e]
        p = e


def reference_index(data: bytes):
    out = []
    for p, line in fgets_lines(data):
        if line[:1] == b">":
            cut = len(line)
            for k in range(len(line) - 1, 0, -1):
                if not (33 <= line[k] <= 126):
                    cut = k
            cut = min(cut, 50)
            out.append((line[1:cut].decode("latin-1").lower(), p + len(line)))
    return out


def reference_load(data: bytes, pos: int) -> bytes:
    got, line_len, keep = bytearray(), None, 0
    for _, line in fgets_lines(data[pos:]):
        if line[:1] == b">":
            break
        if len(got) == 0 or len(line) != line_len:
            line_len = len(line)
            q = len(line) - 1
            while q > 0 and not chr(line[q]).isalpha():
                q -= 1
            keep = q + 1
        got += line[:keep]
    return bytes(got)


CASES = {
    "plain": b">chr1\nACGTACGTAC\nGGGTTTCCCA\nAC\n>chr2 some description\nNNNNACGT\nacgtnnnn\n",
    "crlf": b">c1\r\nACGTACGT\r\nACGTACGT\r\nAC\r\n>c2\r\nGG\r\n",
    "trailing_blanks": b">c1\nACGT  \nGGCC  \nTT \nAAAAAA\n",
    "same_length_other_cut": b">c1\nACGTAC\nACGT  \nACGTAC\n",            # the cut of line 1 is reused for line 2 (same length): blanks kept
    "blank_lines": b">c1\nACGT\n\nACGT\n\n\n>c2\n\nAC\n",
    "no_final_newline": b">c1\nACGTACGT\nACG",
    "long_line": b">c1\n" + b"ACGT" * 700 + b"\n" + b"TTGA" * 300 + b"\n",          # 2800 characters: three fgets lines
    "long_name": b">" + b"Contig_" * 12 + b" descr\nACGT\n>ok\nGG\n",
    "empty_header": b">\nACGT\n> spaced\nGGGG\n>real\nTT\n",
    "tab_in_header": b">chrX\tdesc\nACGTN\n",
    "digits_and_stars": b">c1\nACGT*\nACG12\nAC-GT\n",
    "lowercase_mixed": b">ChrM\nacgtNNacgt\nACGTnnACGT\n",
}


def test_index_and_load_follow_the_reference_rules(tmp_path):
    for tag, data in CASES.items():
        p = tmp_path / f"{tag}.fa"
        p.write_bytes(data)
        want = reference_index(data)
        with hostlib.Fasta(str(p)) as fa:
            assert fa.names == [n for n, _ in want], tag
            for k, (name, pos) in enumerate(want):
                got = fa.load(k).tobytes()
                assert got == reference_load(data, pos), (tag, name)
                if name and [n for n, _ in want].index(name) == k:
                    assert fa.find(name.upper()) == k, tag
            assert fa.find("absent") == -1


def test_generated_reference_equals_the_python_reader(tmp_path):
    spec = synth.SynthSpec(contigs=[("chrA", 70_001), ("chrB", 12_345), ("chrC", 6_060)], depth=1, seed=3)
    cs = synth.simulate(spec)
    path = str(tmp_path / "g.fa")
    synth.write_fasta(path, [(c.name, c.chars) for c in cs])
    py = read_fasta(path)
    with hostlib.Fasta(path) as fa:
        assert fa.names == [c.name.lower() for c in cs]
        for k, c in enumerate(cs):
            got = fa.load(k)
            assert np.array_equal(got, c.chars) and np.array_equal(got, py[c.name])


def test_golden_fasta_of_the_reference_run(tmp_path):
    """The FASTA the committed reference dumps were made from (tests/golden/g1.fa.gz, unpacked): same contigs and characters as the Python
    reader gives for the compressed file -- the two readers of grom_b200.pipeline agree."""
    import gzip
    from util import GOLDEN
    src = os.path.join(GOLDEN, "g1.fa.gz")
    plain = tmp_path / "g1.fa"
    plain.write_bytes(gzip.open(src, "rb").read())
    py = read_fasta(src)
    with hostlib.Fasta(str(plain)) as fa:
        assert fa.names == [k.lower() for k in py] and len(fa.names) >= 3
        for k, name in enumerate(py):
            assert np.array_equal(fa.load(k), py[name])
    from grom_b200.pipeline import _LazyFasta
    lazy = _LazyFasta(str(plain))
    for name in py:
        assert name.lower() in lazy and np.array_equal(lazy[name.lower()], py[name])
    assert "nope" not in lazy
