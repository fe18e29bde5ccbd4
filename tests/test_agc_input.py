"""AGC input of the index builder (pgr-mdb): the reader binds libagc's C API at run time (pgr_tk_b200/host/agc_reader.cpp); the
library itself is built from the reference's vendored agc/ sources when they are present (host/Makefile), so these tests skip
where libagc_ref.so is missing.  Fixture: the reference's own test.agc with the FASTA files it was made from
(pgr-db/test/test_data/gen_agc.sh)."""
import os
import subprocess
import sys

import pytest

import pgr_tk_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "pgr_tk_b200", "host")
LIBAGC = os.path.join(ROOT, "pgr_tk_b200", "libagc_ref.so")
CLI = os.path.join(ROOT, "pgr_tk_b200", "pgr-b200-mdb")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fastx_oracle as fo  # noqa: E402

needs_agc = pytest.mark.skipif(not os.path.exists(LIBAGC), reason="libagc_ref.so not built (needs the reference's agc/ sources)")

PROG = r'''
#include <cstdio>
#include "agc_reader.hpp"
int main(int argc, char **argv) {
    pgrb200::AgcFile f;
    std::string err;
    if (!f.open(argv[1], false, err)) { printf("E %s\n", err.c_str()); return 1; }
    std::vector<std::vector<uint8_t>> seqs;
    if (!f.fetch(0, f.contigs().size(), 3, seqs, err)) { printf("E %s\n", err.c_str()); return 1; }
    for (size_t i = 0; i < seqs.size(); i++) {
        unsigned long long h = 1469598103934665603ull;
        for (unsigned char c : seqs[i]) { h ^= c; h *= 1099511628211ull; }
        printf("R %s %s %zu %llu\n", f.contigs()[i].sample.c_str(), f.contigs()[i].name.c_str(), seqs[i].size(), h);
    }
    return 0;
}
'''


def fnv1a(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def agc_records():
    """(sample, contig, sequence) in archive order: sample test_agc_ref then test_agc_seqs, contigs in file order"""
    out = []
    for sample in ("test_agc_ref", "test_agc_seqs"):
        for name, seq in fo.parse_fasta(open(os.path.join(GOLDEN, sample + ".fa"), "rb").read()):
            out.append((sample, name.decode(), seq))
    return out


@needs_agc
def test_agc_reader_returns_the_fixture_sequences(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(PROG)
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", HOST, "-o", exe, str(src), os.path.join(HOST, "agc_reader.cpp"), "-ldl"])
    env = dict(os.environ, PGR_B200_LIBAGC=LIBAGC)
    txt = subprocess.check_output([exe, os.path.join(GOLDEN, "test.agc")], env=env).decode()
    got = [tuple(ln.split(" ")[1:]) for ln in txt.splitlines() if ln.startswith("R ")]
    exp = [(s, n, str(len(q)), str(fnv1a(q))) for s, n, q in agc_records()]
    assert len(exp) == 66 and sorted(got) == sorted(exp)
    assert [g[0] for g in got] == [e[0] for e in exp]          # samples in archive order


@needs_agc
@pytest.mark.gpu
def test_pgr_mdb_cli_on_the_reference_archive(tmp_path):
    """pgr-b200-mdb on test.agc == the oracle's index of the same sequences with AGC fragment numbering (seq_to_index) and
    sids in archive order; .midx = sid, length, contig, sample"""
    import orc
    recs = {(s, n): q for s, n, q in agc_records()}
    fl = tmp_path / "files.txt"
    fl.write_text(os.path.join(GOLDEN, "test.agc") + "\n")
    prefix = str(tmp_path / "out")
    subprocess.check_call([CLI, str(fl), prefix, "--number-of-readers", "3"], cwd=ROOT)
    midx = [l.rstrip("\n").split("\t") for l in open(prefix + ".midx")]
    assert len(midx) == 66 and [int(r[0]) for r in midx] == list(range(66))
    seqs = [recs[(r[3], r[2])] for r in midx]
    assert [int(r[1]) for r in midx] == [len(s) for s in seqs]
    o = orc.Index(orc.mkspec(80, 56, 4, 64), 1)
    o.add_batch(list(range(66)), seqs)
    o.write_mdb(str(tmp_path / "o.mdb"))
    assert open(prefix + ".mdb", "rb").read() == open(tmp_path / "o.mdb", "rb").read()
    # two archives in the list: sequence ids restart at 0 for the second one (seq_db.rs:543), as the reference does
    fl.write_text((os.path.join(GOLDEN, "test.agc") + "\n") * 2)
    subprocess.check_call([CLI, str(fl), prefix + "2", "--sketch", "-r", "2"], cwd=ROOT)
    midx2 = [l.rstrip("\n").split("\t") for l in open(prefix + "2.midx")]
    assert [int(r[0]) for r in midx2] == list(range(66)) * 2
    o2 = orc.Index(orc.mkspec(80, 56, 2, 64, True), 1)
    o2.add_batch(list(range(66)) * 2, seqs * 2)
    o2.write_mdb(str(tmp_path / "o2.mdb"))
    assert open(prefix + "2.mdb", "rb").read() == open(tmp_path / "o2.mdb", "rb").read()
