"""CPU: the .pdb codec.  The C++ encoder / decoder of the host mirror (pgr_tk_b200/host/pdb_io.hpp, no GPU involved) is
compiled into a small program and checked against the Python restatement of the format (oracle/pbundle_oracle.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pbundle_oracle as pbo  # noqa: E402

PROG = r'''
#include <cstdio>
#include "pdb_io.hpp"
using namespace pgrb200;
int main(int argc, char **argv) {
    PdbData d;
    if (!read_pdb(argv[1], d)) return 3;                 // decode what Python wrote ...
    if (!write_pdb(argv[2], d)) return 4;                // ... and encode it again
    printf("%u %u %u %u %llu %llu %zu %zu\n", d.w, d.k, d.r, d.min_span, (unsigned long long)d.min_branch_size, (unsigned long long)d.min_cov,
           d.bundles.size(), d.vmap.size());
    return 0;
}
'''


def test_pdb_round_trip_between_cpp_and_python(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(PROG)
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "pgr_tk_b200", "host"), "-o", exe, str(src)])
    big = (1 << 56) - 3
    pbid = [(1, 0, [(5, 6, 0), (big, big + 1, 1)]), (0, 2 ** 64 - 1, []), (2, 70000, [(250, 251, 1), (65535, 65536, 0), (2 ** 32, 2 ** 32 + 1, 1)])]
    vmap = {(5, 6): (1, 0, 0), (big, big + 1): (1, 1, 1), (250, 251): (2, 1, 0), (65535, 65536): (2, 0, 300), (2 ** 32, 2 ** 32 + 1): (2, 1, 2)}
    buf = pbo.encode_pdb(48, 56, 4, 12, 8, 0, pbid, vmap)
    assert pbo.decode_pdb(buf) == ((48, 56, 4, 12, 8, 0), pbid, vmap)
    a, b = tmp_path / "a.pdb", tmp_path / "b.pdb"
    a.write_bytes(buf)
    out = subprocess.check_output([exe, str(a), str(b)]).decode().split()
    assert out == ["48", "56", "4", "12", "8", "0", "3", "5"]
    assert b.read_bytes() == buf                           # same bytes: integer widths, map entries by ascending key
    # the varint boundaries
    assert pbo._vi(250) == b"\xfa" and pbo._vi(251) == b"\xfb\xfb\x00" and pbo._vi(65536) == b"\xfc\x00\x00\x01\x00"
