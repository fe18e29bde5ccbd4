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
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, str(src), os.path.join(HOST, "seq_index_db.cpp"), os.path.join(HOST, "fastx_ingest.cpp"), "-pthread",
                           "-L" + libdir, "-lpgr_b200", "-lz", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe, os.path.join(GOLDEN, "test_seqs.fa"), os.path.join(GOLDEN, "test_seqs_frag")]).decode().split()
    assert out == ["66", "66"]
    assert hashlib.md5(open(os.path.join(GOLDEN, "test_seqs_frag.frg"), "rb").read()).hexdigest()   # fixture present


WPROG = r'''
#include <cstdio>
#include "seq_index_db.hpp"
using namespace pgrb200;
// the oracle's fragment compression (test infrastructure) stands in for the GPU call: same record layout
extern "C" int orc_compress_fragments(const pgr_shmmr_spec *spec, size_t n, const uint32_t *sids, const uint8_t *const *seqs, const size_t *lens, int nthreads,
                                      pgr_fragment **frags, size_t *n_frags, pgr_aln_seg **segs, size_t *n_segs);
int main(int argc, char **argv) {
    std::vector<SeqRec> recs;
    std::string err;
    if (!read_fastx(argv[1], recs, err)) return 2;
    std::vector<uint32_t> sids; std::vector<const uint8_t *> ptrs; std::vector<size_t> lens;
    std::vector<CompactSeq> seqs; std::vector<SeqSpan> data;
    for (size_t i = 0; i < recs.size(); i++) {
        sids.push_back((uint32_t)i); ptrs.push_back(recs[i].seq.data()); lens.push_back(recs[i].seq.size());
        CompactSeq cs; cs.id = (uint32_t)i; cs.len = recs[i].seq.size(); cs.name = recs[i].id; cs.source = "test_seqs.fa";
        seqs.push_back(cs); SeqSpan sp; sp.p = recs[i].seq.data(); sp.len = recs[i].seq.size(); data.push_back(sp);
    }
    pgr_shmmr_spec sp{80, 56, 4, 64, 0};
    pgr_fragment *fr = nullptr; pgr_aln_seg *sg = nullptr; size_t nf = 0, ns = 0;
    if (orc_compress_fragments(&sp, recs.size(), sids.data(), ptrs.data(), lens.data(), 2, &fr, &nf, &sg, &ns) != 0) return 3;
    if (write_frag_store(argv[2], 256, 56, fr, nf, sg, seqs, data, err) != 0) { fprintf(stderr, "%s\n", err.c_str()); return 4; }
    printf("%zu %zu\n", nf, ns);
    return 0;
}
'''


def test_frag_store_writer_reproduces_the_reference_store(tmp_path):
    """the product's .sdx/.frg writer (bincode + raw deflate), fed with the oracle's records, against the reference's fixture:
    same .sdx sequence table and chunk base counts, inflated chunk payloads byte-identical"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import frag_format as ff
    import orc
    orc.lib()
    if not os.path.exists(pg.library_path()):
        pg.build_library()
    src = tmp_path / "w.cpp"
    src.write_text(WPROG)
    exe = str(tmp_path / "w")
    libdir, odir = os.path.dirname(pg.library_path()), os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, "-o", exe, str(src), os.path.join(HOST, "seq_index_db.cpp"), os.path.join(HOST, "fastx_ingest.cpp"), "-pthread",
                           "-L" + libdir, "-lpgr_b200", "-L" + odir, "-lpgr_oracle", "-lz", "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir])
    prefix = str(tmp_path / "st")
    out = subprocess.check_output([exe, os.path.join(GOLDEN, "test_seqs.fa"), prefix]).decode().split()
    assert out[0] == "952"
    rcs, raddr, rseqs = ff.read_sdx(os.path.join(GOLDEN, "test_seqs_frag.sdx"))
    gcs, gaddr, gseqs = ff.read_sdx(prefix + ".sdx")
    assert gcs == rcs and gseqs == rseqs and [a[2] for a in gaddr] == [a[2] for a in raddr]
    assert ff.read_frg_chunks(prefix + ".frg", gaddr) == ff.read_frg_chunks(os.path.join(GOLDEN, "test_seqs_frag.frg"), raddr)
