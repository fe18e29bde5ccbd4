"""GPU: the C++ host mirror / CLI (pgr-b200-make-frgdb, same command line as pgr-make-frgdb) writes the reference's
file formats: <prefix>.mdb equals the oracle's canonical .mdb byte for byte and equals the reference's committed fixture
as a map; <prefix>.midx equals the fixture's."""
import os
import subprocess

import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "pgr_tk_b200", "pgr-b200-make-frgdb")


def test_make_frgdb_cli_matches_fixture_and_oracle(tmp_path):
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    fa = os.path.join(GOLDEN, "test_seqs.fa")
    fl = tmp_path / "files.txt"
    fl.write_text(fa + "\n")
    prefix = str(tmp_path / "out")
    subprocess.check_call([CLI, str(fl), prefix], cwd=ROOT)
    _, ref_map, _ = orc.read_mdb_py(os.path.join(GOLDEN, "test_seqs_frag.mdb"))
    spec, got_map, order = orc.read_mdb_py(prefix + ".mdb")
    assert spec == (80, 56, 4, 64, 0) and got_map == ref_map and order == sorted(order)
    o = orc.Index(orc.mkspec(), 0)
    o.load_fasta(fa)
    o.write_mdb(str(tmp_path / "o.mdb"))
    assert open(prefix + ".mdb", "rb").read() == open(tmp_path / "o.mdb", "rb").read()
    ref_midx = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN, "test_seqs_frag.midx"))]
    got_midx = [l.rstrip("\n").split("\t") for l in open(prefix + ".midx")]
    assert len(got_midx) == len(ref_midx) == 66
    for g, r in zip(got_midx, ref_midx):
        assert g[:3] == r[:3] and os.path.basename(g[3]) == os.path.basename(r[3])


def test_make_frgdb_cli_append_and_flags(tmp_path):
    """two files (load_from_fastx then append_from_fastx, ext.rs:152-199): sids and fragment ids continue; non-default spec"""
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    fa1, fa2 = os.path.join(GOLDEN, "test_seqs.fa"), os.path.join(GOLDEN, "test_rev.fa")
    fl = tmp_path / "files.txt"
    fl.write_text(fa1 + "\n" + fa2 + "\n")
    prefix = str(tmp_path / "out2")
    subprocess.check_call([CLI, str(fl), prefix, "-w", "48", "-k", "56", "-r", "4", "--min-span", "12"], cwd=ROOT)
    o = orc.Index(orc.mkspec(48, 56, 4, 12), 0)
    o.load_fasta(fa1)
    o.load_fasta(fa2)
    o.write_mdb(str(tmp_path / "o.mdb"))
    assert open(prefix + ".mdb", "rb").read() == open(tmp_path / "o.mdb", "rb").read()
    assert len(open(prefix + ".midx").readlines()) == 68
    # the .mdb can be loaded back by the library
    assert pg.ShmmrIndex.read_mdb(prefix + ".mdb").as_map() == o.as_map()
