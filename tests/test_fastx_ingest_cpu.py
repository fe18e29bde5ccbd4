"""CPU: the product's FASTA/FASTQ(.gz) ingest (pgr_tk_b200/host/fastx_ingest.cpp: whole-file read, in-place parse, parallel
readers with in-order delivery) against a pure-Python restatement of the reference's readers (oracle/fastx_oracle.py,
fasta_io.rs:46-172) — including the reference's FASTQ quirk of not yielding the last record of a file that ends right after
the quality line — and against the reference's fixture test_seqs.fa."""
import gzip
import os
import subprocess
import sys

import numpy as np
import pytest

import pgr_tk_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fastx_oracle as fo  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "pgr_tk_b200", "host")

PROG = r'''
#include <cstdio>
#include "fastx_ingest.hpp"
using namespace pgrb200;
// argv[1] = number of reader threads, argv[2..] = files; prints "F <path>" then "R <id> <len> <fnv1a of the sequence>" per record
int main(int argc, char **argv) {
    std::vector<std::string> paths;
    for (int i = 2; i < argc; i++) paths.push_back(argv[i]);
    FastxPipeline pipe(paths, atoi(argv[1]), argc > 2 && argv[1][0] == '4' ? INGEST_KEEP : INGEST_PINNED);
    for (size_t i = 0; i < paths.size(); i++) {
        auto pf = pipe.take(i);
        if (!pf->ok) { printf("E %s\n", pf->err.c_str()); continue; }
        printf("F %s\n", pf->path.c_str());
        for (size_t j = 0; j < pf->seqs.size(); j++) {
            unsigned long long h = 1469598103934665603ull;
            for (size_t q = 0; q < pf->seqs[j].len; q++) { h ^= pf->seqs[j].p[q]; h *= 1099511628211ull; }
            printf("R %s %zu %llu\n", pf->ids[j].c_str(), pf->seqs[j].len, h);
        }
    }
    return 0;
}
'''


def fnv1a(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(pg.library_path()):
        pg.build_library()
    d = tmp_path_factory.mktemp("ingest")
    src = d / "t.cpp"
    src.write_text(PROG)
    out = str(d / "t")
    libdir = os.path.dirname(pg.library_path())
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", HOST, "-o", out, str(src), os.path.join(HOST, "fastx_ingest.cpp"),
                           "-L" + libdir, "-lpgr_b200", "-lz", "-Wl,-rpath," + libdir])
    return out


def run(exe, paths, readers=3):
    txt = subprocess.check_output([exe, str(readers)] + paths).decode()
    files, cur = [], None
    for ln in txt.splitlines():
        if ln.startswith("F "):
            cur = []
            files.append(cur)
        elif ln.startswith("E "):
            files.append(ln)
        elif ln.startswith("R "):
            _, rid, n, h = (ln.split(" ") + [""])[:4] if ln.count(" ") == 3 else ("R", "", *ln.split(" ")[-2:])
            cur.append((rid.encode(), int(n), int(h)))
    return files


def expect(buf):
    return [(rid, len(s), fnv1a(s)) for rid, s in fo.parse_fastx(buf)]


def test_fixture_fasta_plain_and_gz(exe, tmp_path):
    fa = os.path.join(GOLDEN, "test_seqs.fa")
    raw = open(fa, "rb").read()
    gz = str(tmp_path / "t.fa.gz")
    with gzip.open(gz, "wb") as f:
        f.write(raw)
    got = run(exe, [fa, gz, fa], readers=3)
    exp = expect(raw)
    assert len(exp) == 66
    assert got == [exp, exp, exp]


CASES = {
    "plain": b">s1 desc here\nACGT\nACG\n>s2\nTTTT\n",
    "crlf": b">s1 d\r\nACGT\r\nAC\r\n>s2\r\nGG\r\n",
    "no_final_newline": b">a\nACGT\n>b\nGGCC",
    "empty_seq_and_blank_lines": b">a\n\n>b\nAC\n\nGT\n\n>c\n",
    "gt_inside_header": b">a>b c\nACGT\n>d\nTT\n",
    "cr_inside_line": b">a\nAC\rGT\n",
    "header_only": b">only\n",
    "fq_trailing_newline": b"@r1 x\nACGT\n+\nIIII\n@r2\nGGCC\n+r2\nJJJJ\n",          # reference drops r2 (fasta_io.rs:160-164)
    "fq_no_trailing_newline": b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\nJJJJ",
    "fq_crlf": b"@r1\r\nAC\r\n+\r\nII\r\n@r2\r\nGT\r\n+\r\nII\r\n@r3\r\nTT\r\n+\r\nII\r\n",
    "fq_one_record": b"@r1\nACGT\n+\nIIII\n",
    "fq_trailing_blank_line": b"@r1\nACGT\n+\nIIII\n@r2\nGGCC\n+\nJJJJ\n\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_record_rules(exe, tmp_path, name):
    p = str(tmp_path / (name + ".fx"))
    open(p, "wb").write(CASES[name])
    got = run(exe, [p], readers=1)
    assert got == [expect(CASES[name])], name


def test_fastq_quirk_is_what_the_reference_does():
    # pinned by reading fasta_io.rs:120-165: `if res.ok() == Some(0) { return None; }` comes BEFORE `Some(Ok(rec))`
    assert [r for r, _ in fo.parse_fastq(CASES["fq_trailing_newline"])] == [b"r1"]
    assert [r for r, _ in fo.parse_fastq(CASES["fq_no_trailing_newline"])] == [b"r1"]
    # only bytes after the last quality line keep the last record (and the empty record that follows is dropped in turn)
    assert [r for r, _ in fo.parse_fastq(CASES["fq_trailing_blank_line"])] == [b"r1", b"r2"]


def test_many_files_in_order_with_random_content(exe, tmp_path):
    rng = np.random.default_rng(3)
    paths, exps = [], []
    for i in range(17):
        recs = []
        for j in range(int(rng.integers(1, 6))):
            L = int(rng.integers(0, 5000))
            s = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=L)].tobytes()
            width = int(rng.integers(1, 120))
            body = b"\n".join(s[a:a + width] for a in range(0, L, width))
            recs.append(b">f%d_r%d some text\n" % (i, j) + body + b"\n")
        buf = b"".join(recs)
        p = str(tmp_path / ("f%02d.fa" % i)) + (".gz" if i % 3 == 0 else "")
        if p.endswith(".gz"):
            with gzip.open(p, "wb") as f:
                f.write(buf)
        else:
            open(p, "wb").write(buf)
        paths.append(p)
        exps.append(expect(buf))
    assert run(exe, paths, readers=4) == exps
