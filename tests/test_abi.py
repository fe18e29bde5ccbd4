"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol include/pgr_b200.h
declares, and compute entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import pgr_tk_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pgr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgr_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(pg.library_path()):
        pg.build_library()
    L = ctypes.CDLL(pg.library_path())
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_pod_layouts():
    assert pg.MM128.itemsize == 16 and pg.SIG.itemsize == 20 and pg.QPAIR.itemsize == 32
    assert pg.HITPAIR.itemsize == 20 and pg.ADJ.itemsize == 40
    assert ctypes.sizeof(pg.ShmmrSpec) == 20
    # PODs of the MAP-graph and fragment entry points (include/pgr_b200.h)
    assert pg.GNODE.itemsize == 24 and pg.DFSNODE.itemsize == 72 and pg.ALNSEG.itemsize == 12 and pg.FRAGMENT.itemsize == 40
    assert pg.FRAGMENT.fields["seg_off"][1] == 32 and pg.DFSNODE.fields["weight"][1] == 52


def test_no_cpu_fallback_without_device():
    if pg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pg.PgrError) as e:
        pg.sequence_to_shmmrs(0, b"ACGT" * 100, pg.ShmmrSpec())
    assert e.value.code == -3
    with pytest.raises(pg.PgrError):
        pg.Ctx(0)


def test_spec_asserts_become_error_codes():
    # shmmrutils.rs:443-445 asserts -> PGR_E_SPEC before any device work
    for bad in (pg.ShmmrSpec(80, 57, 4, 64), pg.ShmmrSpec(129, 56, 4, 64), pg.ShmmrSpec(80, 56, 0, 64), pg.ShmmrSpec(80, 56, 13, 64)):
        with pytest.raises(pg.PgrError) as e:
            pg.sequence_to_shmmrs(0, b"ACGT" * 100, bad)
        assert e.value.code == -2


def test_mdb_map_header_pass_needs_no_device():
    """pgr_b200_mdb_map_open (read_mdb_file_to_frag_locations, seq_db.rs:1409-1471) is host code: the reference's fixture .mdb
    gives its spec, 55 keys and 820 signatures; a file that is no .mdb is refused"""
    golden = os.path.join(ROOT, "tests", "golden")
    m = pg.MdbMap(os.path.join(golden, "test_seqs_frag.mdb"))
    spec, nk, ns = m.info()
    assert (spec.w, spec.k, spec.r, spec.min_span, spec.sketch) == (80, 56, 4, 64, 0)
    assert (nk, ns) == (55, 820)
    m.close()
    with pytest.raises(pg.PgrError):
        pg.MdbMap(os.path.join(golden, "test_seqs.fa"))
