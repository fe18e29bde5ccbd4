"""ctypes binding to the CPU oracle (oracle/libpgr_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "libpgr_oracle.so")

MM128 = np.dtype([("x", "<u8"), ("y", "<u8")])
SIG = np.dtype([("frg_id", "<u4"), ("sid", "<u4"), ("bgn", "<u4"), ("end", "<u4"), ("ori", "u1"), ("pad", "u1", 3)])
QPAIR = np.dtype([("h0", "<u8"), ("h1", "<u8"), ("bgn", "<u4"), ("end", "<u4"), ("ori", "u1"), ("pad", "u1", 7)])
HITPAIR = np.dtype([("qb", "<u4"), ("qe", "<u4"), ("tb", "<u4"), ("te", "<u4"), ("qo", "u1"), ("to", "u1"), ("pad", "u1", 2)])
ADJ = np.dtype([("sid", "<u4"), ("ori0", "u1"), ("ori1", "u1"), ("pad", "u1", 2),
                ("a0", "<u8"), ("a1", "<u8"), ("b0", "<u8"), ("b1", "<u8")])
assert MM128.itemsize == 16 and SIG.itemsize == 20 and QPAIR.itemsize == 32 and HITPAIR.itemsize == 20 and ADJ.itemsize == 40


class Spec(C.Structure):
    _fields_ = [("w", C.c_uint32), ("k", C.c_uint32), ("r", C.c_uint32), ("min_span", C.c_uint32), ("sketch", C.c_uint32)]


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("pgr_oracle.cpp", "frag_oracle.cpp", "pgr_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, sz, u32, u64, i64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int64
        P = C.POINTER
        L.orc_u64hash.restype = u64
        L.orc_u64hash.argtypes = [u64]
        L.orc_free.argtypes = [vp]
        L.orc_sequence_to_shmmrs.argtypes = [u32, vp, sz, P(Spec), C.c_int, P(vp), P(sz)]
        L.orc_shmmrs_batch.argtypes = [sz, vp, vp, vp, P(Spec), C.c_int, C.c_int, P(vp), vp]
        L.orc_index_new.restype = vp
        L.orc_index_new.argtypes = [P(Spec), C.c_int]
        L.orc_index_free.argtypes = [vp]
        L.orc_index_add_batch.argtypes = [vp, sz, vp, vp, vp, C.c_int]
        L.orc_index_load_fasta.argtypes = [vp, C.c_char_p, C.c_int]
        for f in ("orc_index_n_keys", "orc_index_n_sigs", "orc_index_n_seqs"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        L.orc_index_export.argtypes = [vp, vp, vp, vp]
        L.orc_index_write_mdb.argtypes = [vp, C.c_char_p]
        L.orc_index_write_midx.argtypes = [vp, C.c_char_p]
        L.orc_index_read_mdb.restype = vp
        L.orc_index_read_mdb.argtypes = [C.c_char_p]
        L.orc_index_get_spec.argtypes = [vp, P(Spec)]
        L.orc_index_seq_info.argtypes = [vp, sz, P(u32), P(u64), P(C.c_char_p), P(C.c_char_p)]
        L.orc_parse_fasta.argtypes = [C.c_char_p, P(sz), P(vp), P(vp), P(vp)]
        L.orc_raw_query.argtypes = [vp, vp, sz, P(vp), P(sz), P(vp), P(vp)]
        L.orc_sparse_aln.argtypes = [vp, sz, u32, C.c_float, i64, C.c_int, P(sz), P(vp), P(vp), P(vp)]
        L.orc_query_fragment_to_hps.argtypes = [vp, vp, sz, C.c_float, i64, i64, i64, i64, i64, C.c_int,
                                                P(sz), P(vp), P(vp), P(vp), P(vp), P(vp)]
        L.orc_adj_list.argtypes = [vp, sz, vp, sz, C.c_int, P(vp), P(sz)]
        _lib = L
    return _lib


def _take(ptr, n, dtype):
    """copy n items of dtype out of a malloc'd oracle buffer and free it"""
    dtype = np.dtype(dtype)
    if n:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).copy()
    else:
        arr = np.zeros(0, dtype=dtype)
    lib().orc_free(ptr)
    return arr


def _bytes(seq):
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8)
    return np.frombuffer(bytes(seq), dtype=np.uint8)


def u64hash(x):
    return int(lib().orc_u64hash(C.c_uint64(x)))


def mkspec(w=80, k=56, r=4, min_span=64, sketch=False):
    return Spec(w, k, r, min_span, 1 if sketch else 0)


def sequence_to_shmmrs(rid, seq, spec, padding=False):
    a = _bytes(seq)
    out, n = C.c_void_p(), C.c_size_t()
    rc = lib().orc_sequence_to_shmmrs(rid, a.ctypes.data, a.size, C.byref(spec), int(padding), C.byref(out), C.byref(n))
    if rc:
        raise ValueError("oracle rc=%d" % rc)
    return _take(out, n.value, MM128)


def _seq_arrays(seqs):
    arrs = [_bytes(s) for s in seqs]
    n = len(arrs)
    ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * max(1, n))(*[a.size for a in arrs])
    return arrs, ptrs, lens


def shmmrs_batch(rids, seqs, spec, padding=False, nthreads=1):
    arrs, ptrs, lens = _seq_arrays(seqs)
    n = len(arrs)
    r = np.ascontiguousarray(rids, dtype=np.uint32)
    offs = np.zeros(n + 1, dtype=np.uint64)
    out = C.c_void_p()
    rc = lib().orc_shmmrs_batch(n, r.ctypes.data, ptrs, lens, C.byref(spec), int(padding), nthreads, C.byref(out), offs.ctypes.data)
    if rc:
        raise ValueError("oracle rc=%d" % rc)
    return _take(out, int(offs[n]), MM128), offs


class Index:
    """CompactSeqDB index restatement (frag_map + sequence table)"""

    def __init__(self, spec=None, frg_id_mode=0, handle=None):
        self.h = handle if handle is not None else lib().orc_index_new(C.byref(spec), frg_id_mode)
        if not self.h:
            raise ValueError("bad spec / unreadable file")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None

    @classmethod
    def read_mdb(cls, path):
        return cls(handle=lib().orc_index_read_mdb(path.encode()))

    def add_batch(self, sids, seqs, nthreads=1):
        arrs, ptrs, lens = _seq_arrays(seqs)
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        lib().orc_index_add_batch(self.h, len(arrs), s.ctypes.data, ptrs, lens, nthreads)

    def load_fasta(self, path, nthreads=1):
        rc = lib().orc_index_load_fasta(self.h, path.encode(), nthreads)
        if rc:
            raise IOError(path)

    def spec(self):
        s = Spec()
        lib().orc_index_get_spec(self.h, C.byref(s))
        return s

    def export(self):
        nk, ns = lib().orc_index_n_keys(self.h), lib().orc_index_n_sigs(self.h)
        keys = np.zeros((nk, 2), dtype=np.uint64)
        offs = np.zeros(nk + 1, dtype=np.uint64)
        sigs = np.zeros(ns, dtype=SIG)
        lib().orc_index_export(self.h, keys.ctypes.data, offs.ctypes.data, sigs.ctypes.data)
        return keys, offs, sigs

    def as_map(self):
        keys, offs, sigs = self.export()
        m = {}
        for i in range(len(keys)):
            v = sigs[int(offs[i]):int(offs[i + 1])]
            m[(int(keys[i, 0]), int(keys[i, 1]))] = [(int(a["frg_id"]), int(a["sid"]), int(a["bgn"]), int(a["end"]), int(a["ori"])) for a in v]
        return m

    def write_mdb(self, path):
        assert lib().orc_index_write_mdb(self.h, path.encode()) == 0

    def write_midx(self, path):
        assert lib().orc_index_write_midx(self.h, path.encode()) == 0

    def seq_info(self):
        out = []
        for i in range(lib().orc_index_n_seqs(self.h)):
            sid, ln, nm, src = C.c_uint32(), C.c_uint64(), C.c_char_p(), C.c_char_p()
            lib().orc_index_seq_info(self.h, i, C.byref(sid), C.byref(ln), C.byref(nm), C.byref(src))
            out.append((sid.value, ln.value, nm.value.decode(), src.value.decode()))
        return out

    def raw_query(self, seq):
        a = _bytes(seq)
        pairs, n, off, hits = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_void_p()
        lib().orc_raw_query(self.h, a.ctypes.data, a.size, C.byref(pairs), C.byref(n), C.byref(off), C.byref(hits))
        offs = _take(off, n.value + 1, np.uint64)
        return _take(pairs, n.value, QPAIR), offs, _take(hits, int(offs[-1]), SIG)

    def query_fragment_to_hps(self, seq, penalty, max_count=None, max_count_query=None, max_count_target=None,
                              max_aln_span=None, max_gap=None, oriented=False):
        a = _bytes(seq)
        opt = lambda v: -1 if v is None else int(v)
        nt = C.c_size_t()
        sids, tco, sc, cho, ch = (C.c_void_p() for _ in range(5))
        rc = lib().orc_query_fragment_to_hps(self.h, a.ctypes.data, a.size, penalty, opt(max_count), opt(max_count_query),
                                             opt(max_count_target), opt(max_aln_span), opt(max_gap), int(oriented),
                                             C.byref(nt), C.byref(sids), C.byref(tco), C.byref(sc), C.byref(cho), C.byref(ch))
        if rc:
            raise ValueError("oracle rc=%d" % rc)
        n = nt.value
        tco_a = _take(tco, n + 1, np.uint64)
        nch = int(tco_a[-1])
        cho_a = _take(cho, nch + 1, np.uint64)
        return (_take(sids, n, np.uint32), tco_a, _take(sc, nch, np.float32), cho_a, _take(ch, int(cho_a[-1]), HITPAIR))

    def adj_list(self, min_count, keeps=None):
        k = np.ascontiguousarray(keeps if keeps is not None else [], dtype=np.uint32)
        out, n = C.c_void_p(), C.c_size_t()
        lib().orc_adj_list(self.h, min_count, k.ctypes.data, k.size, int(keeps is not None), C.byref(out), C.byref(n))
        return _take(out, n.value, ADJ)


FRAGMENT = np.dtype([("kind", "u1"), ("reversed", "u1"), ("pad", "u1", 2), ("sid", "<u4"), ("bgn", "<u4"), ("end", "<u4"), ("len", "<u4"),
                     ("ref_frag", "<u4"), ("n_segs", "<u4"), ("pad2", "<u4"), ("seg_off", "<u8")])
ALNSEG = np.dtype([("type", "<u4"), ("a", "<u4"), ("b", "<u4")])


def compress_fragments(sids, seqs, spec, nthreads=1):
    """C++ restatement of seq_to_compressed over the sequences in order (oracle/frag_oracle.cpp) -> (FRAGMENT[], ALNSEG[])"""
    arrs, ptrs, lens = _seq_arrays(seqs)
    s = np.ascontiguousarray(sids, dtype=np.uint32)
    fr, sg = C.c_void_p(), C.c_void_p()
    nf, ns = C.c_size_t(), C.c_size_t()
    L = lib()
    L.orc_compress_fragments.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    rc = L.orc_compress_fragments(C.byref(spec), len(arrs), s.ctypes.data, ptrs, lens, nthreads, C.byref(fr), C.byref(nf), C.byref(sg), C.byref(ns))
    if rc:
        raise ValueError("oracle rc=%d" % rc)
    return _take(fr, nf.value, FRAGMENT), _take(sg, ns.value, ALNSEG)


def sparse_aln(hits, max_span, penalty, max_gap=None, oriented=False):
    """hits: HITPAIR array (sorted in place like the reference). Returns (scores, chain_off, chain_hits)."""
    h = np.ascontiguousarray(hits, dtype=HITPAIR)
    nc = C.c_size_t()
    off, sc, ch = C.c_void_p(), C.c_void_p(), C.c_void_p()
    rc = lib().orc_sparse_aln(h.ctypes.data, h.size, max_span, penalty, -1 if max_gap is None else int(max_gap), int(oriented),
                              C.byref(nc), C.byref(off), C.byref(sc), C.byref(ch))
    if rc:
        raise ValueError("oracle rc=%d" % rc)
    off_a = _take(off, nc.value + 1, np.uint64)
    return _take(sc, nc.value, np.float32), off_a, _take(ch, int(off_a[-1]), HITPAIR), h


def parse_fasta(path):
    n = C.c_size_t()
    names, seqs, lens = C.c_void_p(), C.c_void_p(), C.c_void_p()
    rc = lib().orc_parse_fasta(path.encode(), C.byref(n), C.byref(names), C.byref(seqs), C.byref(lens))
    if rc:
        raise IOError(path)
    nn = n.value
    name_ptrs = (C.c_void_p * max(1, nn)).from_address(names.value)
    seq_ptrs = (C.c_void_p * max(1, nn)).from_address(seqs.value)
    len_arr = (C.c_size_t * max(1, nn)).from_address(lens.value)
    out = []
    for i in range(nn):
        nm = C.string_at(name_ptrs[i]).decode()
        s = C.string_at(seq_ptrs[i], len_arr[i]) if len_arr[i] else b""
        lib().orc_free(name_ptrs[i])
        lib().orc_free(seq_ptrs[i])
        out.append((nm, s))
    lib().orc_free(names)
    lib().orc_free(seqs)
    lib().orc_free(lens)
    return out


def read_mdb_py(path):
    """independent pure-Python parser of the .mdb layout (seq_db.rs:1291-1326) used to pin the oracle"""
    import struct
    b = open(path, "rb").read()
    assert b[:3] == b"mdb"
    w, k, r, ms, flag = struct.unpack_from("<5I", b, 3)
    (nk,) = struct.unpack_from("<Q", b, 23)
    c = 31
    m = {}
    order = []
    for _ in range(nk):
        h0, h1, vl = struct.unpack_from("<3Q", b, c)
        c += 24
        v = []
        for _ in range(vl):
            v.append(struct.unpack_from("<4IB", b, c))
            c += 17
        m[(h0, h1)] = v
        order.append((h0, h1))
    assert c == len(b)
    return (w, k, r, ms, flag & 1), m, order
