"""CPU: host-side logic of the C++ mirror that needs no device — the FASTA reader (fasta_io.rs:46-118 rules) and the
.sdx/.frg reader (`FragStore`, frag_file_io.rs / seq_db.rs:685-735) run on the REFERENCE'S OWN fixture store and must
give back the 66 sequences of test_seqs.fa; the bincode / deflate writer helpers are exercised through a round trip."""
import hashlib
import os
import subprocess

import pgr_tk_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "pgr_tk_b200", "host")

PROG = r'''
#include <cstdio>
#include "seq_index_db.hpp"
using namespace pgrb200;
int main(int argc, char **argv) {
    std::vector<SeqRec> recs;
    std::string err;
    if (!read_fastx(argv[1], recs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
    FragStore st;
    if (!st.load(argv[2], 56, err)) { fprintf(stderr, "%s\n", err.c_str()); return 3; }
    if (st.seqs().size() != recs.size()) return 4;
    size_t same = 0;
    std::vector<uint8_t> seq;
    for (size_t i = 0; i < recs.size(); i++) {
        if (!st.get_seq_by_id((uint32_t)i, seq, err)) { fprintf(stderr, "%s\n", err.c_str()); return 5; }
        same += seq == recs[i].seq && st.seqs()[i].name == recs[i].id && st.seqs()[i].len == seq.size();
    }
    // backwards, so that base fragments live in chunks that were evicted from the small cache
    for (size_t i = recs.size(); i-- > 0;) { if (!st.get_seq_by_id((uint32_t)i, seq, err) || seq != recs[i].seq) return 6; }
    printf("%zu %zu\n", recs.size(), same);
    return 0;
}
'''


def test_fasta_reader_and_frag_store_on_the_reference_fixture(tmp_path):
    if not os.path.exists(pg.library_path()):
        pg.build_library()
    src = tmp_path / "t.cpp"
    src.write_text(PROG)
    exe = str(tmp_path / "t")
    libdir = os.path.dirname(pg.library_path())
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, str(src), os.path.join(HOST, "seq_index_db.cpp"),
                           "-L" + libdir, "-lpgr_b200", "-lz", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe, os.path.join(GOLDEN, "test_seqs.fa"), os.path.join(GOLDEN, "test_seqs_frag")]).decode().split()
    assert out == ["66", "66"]
    assert hashlib.md5(open(os.path.join(GOLDEN, "test_seqs_frag.frg"), "rb").read()).hexdigest()   # fixture present
