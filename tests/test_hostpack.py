"""CPU: the host half of the packed transport (pgr_b200_pack_bases, hostpack.cpp) against a numpy restatement of the
reference's byte classes (shmmrutils.rs:426-436), on every SIMD path this machine has, every length modulo 64, every byte
value."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pgr_tk_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def planes_expected(a):
    """(p0, p1, v) per 32-byte block from the LUT: bytes 0..3 and ACGTacgt are bases, anything else is not; the padding of the
    last block counts as a base with code 0"""
    lut = np.full(256, 4, dtype=np.uint8)
    for ch, c in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3)):
        lut[ch[0]] = c
        lut[ch.lower()[0]] = c
    lut[:4] = np.arange(4)
    code = lut[a]
    nb = (len(a) + 31) // 32
    full = np.full(nb * 32, 0, dtype=np.uint8)   # padding = code 0
    full[:len(a)] = code
    full = full.reshape(nb, 32)
    sh = np.arange(32, dtype=np.uint64)
    def plane(mask):
        return (mask.astype(np.uint64) << sh).sum(axis=1).astype(np.uint32)
    ok = full < 4
    return plane(((full & 1) == 1) & ok), plane(((full >> 1) & 1 == 1) & ok), plane(ok)


def check_all(rng):
    cases = [np.arange(256, dtype=np.uint8).repeat(3), np.zeros(0, dtype=np.uint8)]
    for L in list(range(0, 200)) + [1000, 4097, 65536 + 17]:
        alphabet = np.frombuffer(b"ACGTacgtNn\x00\x01\x02\x03\x04-*Xx", dtype=np.uint8)
        cases.append(alphabet[rng.integers(0, len(alphabet), size=L)])
        cases.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)])
        cases.append(rng.integers(0, 256, size=L, dtype=np.uint8))
    for a in cases:
        for shift in (0, 1, 7):   # unaligned sources
            buf = np.zeros(len(a) + shift, dtype=np.uint8)
            buf[shift:] = a
            got = pg.pack_bases(buf[shift:])
            exp = planes_expected(a)
            for g, e, name in zip(got, exp, ("p0", "p1", "v")):
                assert np.array_equal(g, e), (pg.pack_isa(), name, len(a), shift)
            # the return value: AND of the validity words (what lets a slot leave its validity plane at home)
            exp_all = int(np.bitwise_and.reduce(exp[2])) if len(exp[2]) else 0xFFFFFFFF
            assert pg.pack_bases.last_all_valid == exp_all, (pg.pack_isa(), len(a), shift)


def test_pack_bases_matches_the_lut():
    check_all(np.random.default_rng(5))


@pytest.mark.parametrize("isa", ["scalar", "avx2"])
def test_pack_bases_other_isas(isa):
    # the SIMD path is picked once per process: run the same check in a child with the path forced
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np, pgr_tk_b200 as pg, test_hostpack as t; "
            "t.check_all(np.random.default_rng(6)); print(pg.pack_isa())" % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, PGR_B200_PACK_ISA=isa)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() in (isa, "scalar")   # a machine without AVX2 falls back to scalar
