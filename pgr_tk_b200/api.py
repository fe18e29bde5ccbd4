"""ctypes binding of libpgr_b200.so (C ABI: include/pgr_b200.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGR_B200_LIB points at an alternative build of the same library (tuning aid)
_LIB = os.environ.get("PGR_B200_LIB") or os.path.join(_HERE, "libpgr_b200.so")

MM128 = np.dtype([("x", "<u8"), ("y", "<u8")])
SIG = np.dtype([("frg_id", "<u4"), ("sid", "<u4"), ("bgn", "<u4"), ("end", "<u4"), ("ori", "u1"), ("pad", "u1", 3)])
QPAIR = np.dtype([("h0", "<u8"), ("h1", "<u8"), ("bgn", "<u4"), ("end", "<u4"), ("ori", "u1"), ("pad", "u1", 7)])
HITPAIR = np.dtype([("qb", "<u4"), ("qe", "<u4"), ("tb", "<u4"), ("te", "<u4"), ("qo", "u1"), ("to", "u1"), ("pad", "u1", 2)])
GNODE = np.dtype([("h0", "<u8"), ("h1", "<u8"), ("ori", "u1"), ("pad", "u1", 7)])
DFSNODE = np.dtype([("node", GNODE), ("prev", GNODE), ("has_prev", "u1"), ("is_leaf", "u1"), ("pad", "u1", 2),
                    ("weight", "<u4"), ("rank", "<u4"), ("branch", "<u4"), ("branch_rank", "<u4"), ("pad2", "<u4")])
assert GNODE.itemsize == 24 and DFSNODE.itemsize == 72
ALNSEG = np.dtype([("type", "<u4"), ("a", "<u4"), ("b", "<u4")])
FRAGMENT = np.dtype([("kind", "u1"), ("reversed", "u1"), ("pad", "u1", 2), ("sid", "<u4"), ("bgn", "<u4"), ("end", "<u4"), ("len", "<u4"),
                     ("ref_frag", "<u4"), ("n_segs", "<u4"), ("pad2", "<u4"), ("seg_off", "<u8")])
assert ALNSEG.itemsize == 12 and FRAGMENT.itemsize == 40
ADJ = np.dtype([("sid", "<u4"), ("ori0", "u1"), ("ori1", "u1"), ("pad", "u1", 2),
                ("a0", "<u8"), ("a1", "<u8"), ("b0", "<u8"), ("b1", "<u8")])


class PgrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pgr_b200 error %d: %s" % (code, msg))
        self.code = code


class ShmmrSpec(C.Structure):
    """shmmrutils.rs:20-27"""
    _fields_ = [("w", C.c_uint32), ("k", C.c_uint32), ("r", C.c_uint32), ("min_span", C.c_uint32), ("sketch", C.c_uint32)]

    def __init__(self, w=80, k=56, r=4, min_span=64, sketch=False):
        super().__init__(w, k, r, min_span, 1 if sketch else 0)


class QueryParams(C.Structure):
    """arguments of query_fragment_to_hps (aln.rs:147-158); Option<u32> = negative for None"""
    _fields_ = [("penalty", C.c_float), ("max_count", C.c_int64), ("max_count_query", C.c_int64), ("max_count_target", C.c_int64),
                ("max_aln_span", C.c_int64), ("max_gap", C.c_int64), ("oriented", C.c_int32)]


class ShardStats(C.Structure):
    """pgr_shard_stats: what one rank did in the multi-GPU build"""
    _fields_ = [("rank", C.c_uint32), ("n_ranks", C.c_uint32), ("n_tuples_local", C.c_uint64), ("n_tuples_sent", C.c_uint64),
                ("n_tuples_owned", C.c_uint64), ("bytes_sent", C.c_uint64), ("bytes_recv", C.c_uint64), ("total_frags", C.c_uint64),
                ("stage_ms", C.c_float), ("partition_ms", C.c_float), ("exchange_ms", C.c_float), ("sort_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class QueryResult(C.Structure):
    _fields_ = [("n_queries", C.c_size_t), ("n_targets", C.c_size_t), ("n_chains", C.c_size_t), ("n_hits", C.c_size_t),
                ("q_target_off", C.POINTER(C.c_uint64)), ("target_sid", C.POINTER(C.c_uint32)),
                ("target_chain_off", C.POINTER(C.c_uint64)), ("chain_score", C.POINTER(C.c_float)),
                ("chain_hit_off", C.POINTER(C.c_uint64)), ("hits", C.c_void_p)]


def library_path():
    return _LIB


def build_library(force=False):
    """compile libpgr_b200.so in-tree (nvcc, sm_100a)"""
    csrc = os.path.join(_HERE, "csrc")
    if force and os.path.exists(_LIB):
        os.remove(_LIB)
    subprocess.check_call(["make", "-C", csrc, "-s"])
    return _LIB


_lib = None


def _preload_nccl():
    """libpgr_b200.so links libnccl.so.2 by soname.  When PyTorch's bundled NCCL is installed, map that copy first so that
    the process holds ONE NCCL whatever the import order (torch after pgr_tk_b200 would otherwise get the system copy)."""
    try:
        import importlib.util
        sp = importlib.util.find_spec("nvidia.nccl")
        for d in (sp.submodule_search_locations if sp else []):
            path = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return path
    except Exception:
        pass
    return None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise PgrError(-3, "libpgr_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        _preload_nccl()
        L = C.CDLL(_LIB)
        vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
        P = C.POINTER
        L.pgr_b200_device_count.restype = C.c_int
        L.pgr_b200_last_error.restype = C.c_char_p
        L.pgr_b200_set_default_device.argtypes = [C.c_int]
        L.pgr_b200_free.argtypes = [vp]
        L.pgr_b200_host_alloc.restype = vp
        L.pgr_b200_host_alloc.argtypes = [sz]
        L.pgr_b200_host_free.argtypes = [vp]
        L.pgr_b200_sequence_to_shmmrs.argtypes = [u32, vp, sz, P(ShmmrSpec), C.c_int, P(vp), P(sz)]
        L.pgr_b200_shmmrs_batch.argtypes = [sz, vp, vp, vp, P(ShmmrSpec), C.c_int, P(vp), vp]
        L.pgr_b200_ctx_new.restype = vp
        L.pgr_b200_ctx_new.argtypes = [C.c_int]
        L.pgr_b200_ctx_free.argtypes = [vp]
        L.pgr_b200_ctx_set_stream.argtypes = [vp, vp]
        L.pgr_b200_ctx_upload.argtypes = [vp, sz, vp, vp, vp]
        L.pgr_b200_ctx_set_device_seqs.argtypes = [vp, vp, sz, vp, vp, vp]
        L.pgr_b200_ctx_shmmrs.argtypes = [vp, P(ShmmrSpec), C.c_int, P(sz)]
        L.pgr_b200_ctx_shmmrs_device.argtypes = [vp, P(vp), P(vp)]
        L.pgr_b200_ctx_shmmrs_download.argtypes = [vp, P(vp), vp]
        L.pgr_b200_ctx_timings.argtypes = [vp, P(P(C.c_char_p)), P(P(C.c_float)), P(sz)]
        L.pgr_b200_ctx_counters.argtypes = [vp, P(u64 * 8)]
        L.pgr_b200_index_new.restype = vp
        L.pgr_b200_index_new.argtypes = [P(ShmmrSpec), C.c_int, C.c_int]
        L.pgr_b200_index_free.argtypes = [vp]
        L.pgr_b200_index_get_spec.argtypes = [vp, P(ShmmrSpec)]
        L.pgr_b200_index_add_batch.argtypes = [vp, sz, vp, vp, vp]
        L.pgr_b200_index_stage_batch.argtypes = [vp, sz, vp, vp, vp, P(u64)]
        L.pgr_b200_index_commit_batch.argtypes = [vp, u32]
        L.pgr_b200_index_stage_device.argtypes = [vp, vp, sz, vp, vp, vp, P(u64)]
        L.pgr_b200_index_finalize.argtypes = [vp]
        L.pgr_b200_comm_unique_id.argtypes = [vp]
        L.pgr_b200_comm_init_rank.restype = vp
        L.pgr_b200_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.pgr_b200_comm_free.argtypes = [vp]
        L.pgr_b200_index_build_sharded.argtypes = [vp, vp, sz, vp, vp, vp, P(ShardStats)]
        L.pgr_b200_index_build_sharded_device.argtypes = [vp, vp, vp, sz, vp, vp, vp, P(ShardStats)]
        L.pgr_b200_index_merge.argtypes = [vp, vp, P(ShardStats)]
        L.pgr_b200_mindex_new.restype = vp
        L.pgr_b200_mindex_new.argtypes = [P(ShmmrSpec), C.c_int, C.c_int]
        L.pgr_b200_mindex_new_devices.restype = vp
        L.pgr_b200_mindex_new_devices.argtypes = [P(ShmmrSpec), C.c_int, C.c_int, vp]
        L.pgr_b200_mindex_free.argtypes = [vp]
        L.pgr_b200_mindex_n_shards.argtypes = [vp]
        L.pgr_b200_mindex_add_batch.argtypes = [vp, sz, vp, vp, vp]
        L.pgr_b200_mindex_finalize.argtypes = [vp]
        L.pgr_b200_mindex_counts.argtypes = [vp, P(sz), P(sz), P(u32)]
        L.pgr_b200_mindex_stats.argtypes = [vp, C.c_int, P(ShardStats)]
        L.pgr_b200_mindex_shard.restype = vp
        L.pgr_b200_mindex_shard.argtypes = [vp, C.c_int]
        L.pgr_b200_mindex_export_csr.argtypes = [vp, vp, vp, vp]
        L.pgr_b200_mindex_write_mdb.argtypes = [vp, C.c_char_p]
        L.pgr_b200_mindex_gather.restype = vp
        L.pgr_b200_mindex_gather.argtypes = [vp, C.c_int]
        L.pgr_b200_host_register.argtypes = [vp, sz]
        L.pgr_b200_host_unregister.argtypes = [vp]
        L.pgr_b200_pack_bases.restype = C.c_uint32
        L.pgr_b200_pack_bases.argtypes = [vp, sz, vp, vp, vp]
        L.pgr_b200_pack_isa.restype = C.c_char_p
        L.pgr_b200_transport_bytes.restype = C.c_uint64
        L.pgr_b200_index_counts.argtypes = [vp, P(sz), P(sz), P(u32)]
        L.pgr_b200_index_export_csr.argtypes = [vp, vp, vp, vp]
        L.pgr_b200_index_tuples_device.argtypes = [vp, P(vp), P(sz)]
        L.pgr_b200_index_set_tuples_device.argtypes = [vp, vp, sz]
        L.pgr_b200_index_partition.argtypes = [vp, sz, vp, vp]
        L.pgr_b200_index_write_mdb.argtypes = [vp, C.c_char_p]
        L.pgr_b200_index_read_mdb.restype = vp
        L.pgr_b200_index_read_mdb.argtypes = [C.c_char_p, C.c_int]
        L.pgr_b200_raw_query.argtypes = [vp, vp, sz, P(vp), P(sz), P(vp), P(vp)]
        L.pgr_b200_query_batch.argtypes = [vp, sz, vp, vp, P(QueryParams), P(P(QueryResult))]
        L.pgr_b200_query_result_free.argtypes = [P(QueryResult)]
        L.pgr_b200_sparse_aln.argtypes = [vp, sz, u32, C.c_float, C.c_int64, C.c_int, P(sz), P(vp), P(vp), P(vp)]
        L.pgr_b200_adj_list.argtypes = [vp, sz, vp, sz, C.c_int, P(vp), P(sz)]
        L.pgr_b200_smp_adj_list_for_seqs.argtypes = [vp, sz, vp, vp, vp, vp, P(vp), P(sz)]
        L.pgr_b200_index_compress_fragments.argtypes = [vp, sz, vp, vp, vp, P(vp), P(sz), P(vp), P(sz)]
        L.pgr_b200_sort_adj_list_by_weighted_dfs.argtypes = [vp, vp, sz, vp, P(vp), P(sz)]
        L.pgr_b200_principal_bundles.argtypes = [vp, vp, sz, sz, P(vp), P(vp), P(sz), P(vp), P(sz)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise PgrError(rc, lib().pgr_b200_last_error().decode(errors="replace"))


def device_count():
    return lib().pgr_b200_device_count()


def set_default_device(device):
    _check(lib().pgr_b200_set_default_device(device))


def _bytes(seq):
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8)
    return np.frombuffer(bytes(seq), dtype=np.uint8)


def _take(ptr, n, dtype):
    dtype = np.dtype(dtype)
    if n:
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).copy()
    else:
        arr = np.zeros(0, dtype=dtype)
    if ptr.value:
        lib().pgr_b200_free(ptr)
    return arr


def _seq_arrays(seqs):
    arrs = [_bytes(s) for s in seqs]
    n = len(arrs)
    ptrs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in arrs])
    lens = (C.c_size_t * max(1, n))(*[a.size for a in arrs])
    return arrs, ptrs, lens


def sequence_to_shmmrs(rid, seq, spec, padding=False):
    """shmmrutils::sequence_to_shmmrs(rid, &seq, &spec, padding) -> Vec<MM128>  (shmmrutils.rs:657-669)"""
    a = _bytes(seq)
    out, n = C.c_void_p(), C.c_size_t()
    _check(lib().pgr_b200_sequence_to_shmmrs(rid, a.ctypes.data, a.size, C.byref(spec), int(padding), C.byref(out), C.byref(n)))
    return _take(out, n.value, MM128)


def get_shmmrs_from_seqs(rids, seqs, spec, padding=False):
    """CompactSeqDB::get_shmmrs_from_seqs (seq_db.rs:456-469) -> (MM128[total], offsets[n+1])"""
    arrs, ptrs, lens = _seq_arrays(seqs)
    n = len(arrs)
    r = np.ascontiguousarray(rids, dtype=np.uint32)
    offs = np.zeros(n + 1, dtype=np.uint64)
    out = C.c_void_p()
    _check(lib().pgr_b200_shmmrs_batch(n, r.ctypes.data, ptrs, lens, C.byref(spec), int(padding), C.byref(out), offs.ctypes.data))
    return _take(out, int(offs[n]), MM128), offs


def pack_bases(seq):
    """host half of the packed transport (pgr_b200_pack_bases): the three bit planes (p0, p1, v) of `seq`, one u32 per
    32-byte block each; runs on the host, no device needed"""
    a = np.frombuffer(_bytes(seq), dtype=np.uint8)
    nb = (len(a) + 31) // 32
    p0, p1, v = (np.zeros(max(nb, 1), dtype=np.uint32) for _ in range(3))
    all_valid = lib().pgr_b200_pack_bases(a.ctypes.data, len(a), p0.ctypes.data, p1.ctypes.data, v.ctypes.data)
    pack_bases.last_all_valid = int(all_valid) & 0xFFFFFFFF   # AND of the validity words (all ones: bases only)
    return p0[:nb], p1[:nb], v[:nb]


TRANSPORT_PACKED, TRANSPORT_DIRECT = 0, 1


def set_transport(mode):
    """process-wide host-to-device transport of large inputs (pgr_b200_set_transport); returns the previous mode"""
    return lib().pgr_b200_set_transport(int(mode))


def pack_isa():
    return lib().pgr_b200_pack_isa().decode()


class HostBuffer:
    """pinned host memory (pgr_b200_host_alloc) exposed as a numpy uint8 array"""

    def __init__(self, nbytes):
        self.ptr = lib().pgr_b200_host_alloc(nbytes)
        if not self.ptr:
            _check(-4)
        self.nbytes = nbytes
        self.array = np.frombuffer((C.c_uint8 * nbytes).from_address(self.ptr), dtype=np.uint8)

    def free(self):
        if self.ptr:
            self.array = None
            lib().pgr_b200_host_free(self.ptr)
            self.ptr = None


def host_alloc(nbytes):
    return HostBuffer(nbytes)


class Ctx:
    """explicit device context (one per device / thread)"""

    def __init__(self, device=0):
        self.h = lib().pgr_b200_ctx_new(device)
        if not self.h:
            _check(-3)
        self.n_seq = 0

    def close(self):
        if getattr(self, "h", None):
            lib().pgr_b200_ctx_free(self.h)
            self.h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        _check(lib().pgr_b200_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def upload(self, seqs, rids=None):
        arrs, ptrs, lens = _seq_arrays(seqs)
        n = len(arrs)
        r = np.ascontiguousarray(rids if rids is not None else np.arange(n), dtype=np.uint32)
        _check(lib().pgr_b200_ctx_upload(self.h, n, r.ctypes.data, ptrs, lens))
        self.n_seq = n

    def upload_ptrs(self, ptrs, lens, rids=None):
        """upload from raw host pointers (e.g. slices of a pinned HostBuffer)"""
        n = len(ptrs)
        p = (C.c_void_p * max(1, n))(*ptrs)
        l = (C.c_size_t * max(1, n))(*lens)
        r = np.ascontiguousarray(rids if rids is not None else np.arange(n), dtype=np.uint32)
        _check(lib().pgr_b200_ctx_upload(self.h, n, r.ctypes.data, p, l))
        self.n_seq = n

    def set_device_seqs(self, dev_base, offs, lens, rids=None):
        n = len(offs)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        l = np.ascontiguousarray(lens, dtype=np.uint64)
        r = np.ascontiguousarray(rids if rids is not None else np.arange(n), dtype=np.uint32)
        _check(lib().pgr_b200_ctx_set_device_seqs(self.h, C.c_void_p(dev_base), n, r.ctypes.data, o.ctypes.data, l.ctypes.data))
        self.n_seq = n

    def shmmrs(self, spec, padding=False):
        n = C.c_size_t()
        _check(lib().pgr_b200_ctx_shmmrs(self.h, C.byref(spec), int(padding), C.byref(n)))
        return n.value

    def shmmrs_download(self):
        offs = np.zeros(self.n_seq + 1, dtype=np.uint64)
        out = C.c_void_p()
        _check(lib().pgr_b200_ctx_shmmrs_download(self.h, C.byref(out), offs.ctypes.data))
        return _take(out, int(offs[self.n_seq]), MM128), offs

    def timings(self):
        names, ms, n = C.POINTER(C.c_char_p)(), C.POINTER(C.c_float)(), C.c_size_t()
        _check(lib().pgr_b200_ctx_timings(self.h, C.byref(names), C.byref(ms), C.byref(n)))
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def counters(self):
        out = (C.c_uint64 * 8)()
        _check(lib().pgr_b200_ctx_counters(self.h, C.byref(out)))
        return list(out)


TUPLE = np.dtype([("h0", "<u8"), ("h1", "<u8"), ("frg_id", "<u4"), ("sid", "<u4"), ("bgn", "<u4"), ("end", "<u4"),
                  ("ori", "<u4"), ("pad", "<u4")])

FRG_ID_FASTX, FRG_ID_AGC = 0, 1


class ShmmrIndex:
    """device-resident ShmmrFragMap: the index part of CompactSeqDB (seq_db.rs:95-100)"""

    def __init__(self, spec=None, frg_id_mode=FRG_ID_FASTX, device=-1, handle=None):
        self.h = handle if handle is not None else lib().pgr_b200_index_new(C.byref(spec), frg_id_mode, device)
        if not self.h:
            _check(-3 if device_count() == 0 else -1)

    def close(self):
        if getattr(self, "h", None):
            lib().pgr_b200_index_free(self.h)
            self.h = None

    __del__ = close

    @classmethod
    def read_mdb(cls, path, device=-1):
        """read_mdb_file (seq_db.rs:1328-1407)"""
        h = lib().pgr_b200_index_read_mdb(path.encode(), device)
        if not h:
            _check(-5)
        return cls(handle=h)

    def spec(self):
        s = ShmmrSpec()
        _check(lib().pgr_b200_index_get_spec(self.h, C.byref(s)))
        return s

    def add_batch(self, sids, seqs):
        """load_seqs_from_seq_vec / load_index_from_seq_vec, index part (seq_db.rs:507-525, :573-615)"""
        arrs, ptrs, lens = _seq_arrays(seqs)
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        _check(lib().pgr_b200_index_add_batch(self.h, len(arrs), s.ctypes.data, ptrs, lens))

    def stage_batch(self, sids, seqs):
        arrs, ptrs, lens = _seq_arrays(seqs)
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        nf = C.c_uint64()
        _check(lib().pgr_b200_index_stage_batch(self.h, len(arrs), s.ctypes.data, ptrs, lens, C.byref(nf)))
        return nf.value

    def commit_batch(self, frag_base):
        _check(lib().pgr_b200_index_commit_batch(self.h, frag_base))

    def finalize(self):
        _check(lib().pgr_b200_index_finalize(self.h))

    def build_sharded(self, comm, sids, seqs):
        """collective over `comm` (multi-GPU build, one process per GPU): this rank's block of HOST sequences in, this
        rank's key range (finalized) out.  Returns the rank's pgr_shard_stats as a dict."""
        arrs, ptrs, lens = _seq_arrays(seqs)
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        st = ShardStats()
        _check(lib().pgr_b200_index_build_sharded(self.h, comm.h, len(arrs), s.ctypes.data, ptrs, lens, C.byref(st)))
        return st.as_dict()

    def build_sharded_ptrs(self, comm, sids, ptrs, lens):
        """build_sharded from raw host pointers (slices of a pinned HostBuffer)"""
        n = len(ptrs)
        p = (C.c_void_p * max(1, n))(*ptrs)
        l = (C.c_size_t * max(1, n))(*lens)
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        st = ShardStats()
        _check(lib().pgr_b200_index_build_sharded(self.h, comm.h, n, s.ctypes.data, p, l, C.byref(st)))
        return st.as_dict()

    def build_sharded_device(self, comm, dev_base, sids, offs, lens):
        """build_sharded with the block already resident in this rank's HBM"""
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        l = np.ascontiguousarray(lens, dtype=np.uint64)
        st = ShardStats()
        _check(lib().pgr_b200_index_build_sharded_device(self.h, comm.h, C.c_void_p(dev_base), len(s), s.ctypes.data, o.ctypes.data, l.ctypes.data, C.byref(st)))
        return st.as_dict()

    def stage_device(self, dev_base, sids, offs, lens):
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        o = np.ascontiguousarray(offs, dtype=np.uint64)
        l = np.ascontiguousarray(lens, dtype=np.uint64)
        nf = C.c_uint64()
        _check(lib().pgr_b200_index_stage_device(self.h, C.c_void_p(dev_base), len(s), s.ctypes.data, o.ctypes.data, l.ctypes.data, C.byref(nf)))
        return nf.value

    def counts(self):
        nk, ns, nf = C.c_size_t(), C.c_size_t(), C.c_uint32()
        _check(lib().pgr_b200_index_counts(self.h, C.byref(nk), C.byref(ns), C.byref(nf)))
        return nk.value, ns.value, nf.value

    def export(self):
        """canonical CSR: keys[n_keys,2] ascending, offsets[n_keys+1], sigs[n_sigs]"""
        nk, ns, _ = self.counts()
        keys = np.zeros((max(nk, 1), 2), dtype=np.uint64)
        offs = np.zeros(nk + 1, dtype=np.uint64)
        sigs = np.zeros(max(ns, 1), dtype=SIG)
        _check(lib().pgr_b200_index_export_csr(self.h, keys.ctypes.data, offs.ctypes.data, sigs.ctypes.data))
        return keys[:nk], offs, sigs[:ns]

    def as_map(self):
        keys, offs, sigs = self.export()
        m = {}
        for i in range(len(keys)):
            v = sigs[int(offs[i]):int(offs[i + 1])]
            m[(int(keys[i, 0]), int(keys[i, 1]))] = [(int(a["frg_id"]), int(a["sid"]), int(a["bgn"]), int(a["end"]), int(a["ori"])) for a in v]
        return m

    def tuples_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().pgr_b200_index_tuples_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_tuples_device(self, dev_ptr, n):
        _check(lib().pgr_b200_index_set_tuples_device(self.h, C.c_void_p(dev_ptr), n))

    def partition(self, splitters):
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        counts = np.zeros(len(sp) + 1, dtype=np.uint64)
        _check(lib().pgr_b200_index_partition(self.h, len(sp) + 1, sp.ctypes.data, counts.ctypes.data))
        return counts

    def write_mdb(self, path):
        """write_shmmr_map_file (seq_db.rs:1291-1326), keys ascending"""
        _check(lib().pgr_b200_index_write_mdb(self.h, path.encode()))

    def raw_query(self, seq):
        """raw_query_fragment (seq_db.rs:1200-1228) -> (pairs QPAIR[n], hit_off[n+1], hits SIG[...])"""
        a = _bytes(seq)
        pairs, n, off, hits = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_void_p()
        _check(lib().pgr_b200_raw_query(self.h, a.ctypes.data, a.size, C.byref(pairs), C.byref(n), C.byref(off), C.byref(hits)))
        offs = _take(off, n.value + 1, np.uint64)
        return _take(pairs, n.value, QPAIR), offs, _take(hits, int(offs[-1]), SIG)

    def query_batch(self, seqs, penalty, max_count=None, max_count_query=None, max_count_target=None, max_aln_span=None,
                    max_gap=None, oriented=False):
        """query_fragment_to_hps (ext.rs:252-282) for every sequence of `seqs`.
        Returns (q_target_off, target_sid, target_chain_off, chain_score, chain_hit_off, hits HITPAIR[...])"""
        arrs, ptrs, lens = _seq_arrays(seqs)
        opt = lambda v: -1 if v is None else int(v)
        prm = QueryParams(penalty, opt(max_count), opt(max_count_query), opt(max_count_target), opt(max_aln_span), opt(max_gap), int(oriented))
        res = C.POINTER(QueryResult)()
        _check(lib().pgr_b200_query_batch(self.h, len(arrs), ptrs, lens, C.byref(prm), C.byref(res)))
        return _take_query_result(res)

    def query_fragment_to_hps(self, seq, penalty, **kw):
        """single-query form: (target_sid, target_chain_off, chain_score, chain_hit_off, hits)"""
        r = self.query_batch([seq], penalty, **kw)
        return r[1:]

    def adj_list(self, min_count, keeps=None):
        """frag_map_to_adj_list (seq_db.rs:876-944) -> ADJ[...]"""
        k = np.ascontiguousarray(keeps if keeps is not None else [], dtype=np.uint32)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().pgr_b200_adj_list(self.h, min_count, k.ctypes.data, k.size, int(keeps is not None), C.byref(out), C.byref(n)))
        return _take(out, n.value, ADJ)

    def smp_adj_list_for_seqs(self, sids, seqs, min_counts):
        """generate_smp_adj_list_for_seq (seq_db.rs:946-1000) for every (sid, seq, min_count) -> ADJ[...] in the order given"""
        keep, ptrs, lens = _seq_arrays(seqs)
        sid_a = np.ascontiguousarray(sids, dtype=np.uint32)
        mc = np.ascontiguousarray(min_counts, dtype=np.uint64)
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().pgr_b200_smp_adj_list_for_seqs(self.h, len(keep), sid_a.ctypes.data, ptrs, lens, mc.ctypes.data, C.byref(out), C.byref(n)))
        del keep
        return _take(out, n.value, ADJ)

    def compress_fragments(self, sids, seqs):
        """CompactSeqDB.frags after load_seqs (seq_db.rs:189-358) -> (FRAGMENT[n_frags], ALNSEG[n_segs])"""
        keep, ptrs, lens = _seq_arrays(seqs)
        sid_a = np.ascontiguousarray(sids, dtype=np.uint32)
        fr, sg = C.c_void_p(), C.c_void_p()
        nf, ns = C.c_size_t(), C.c_size_t()
        _check(lib().pgr_b200_index_compress_fragments(self.h, len(seqs), sid_a.ctypes.data, ptrs, lens, C.byref(fr), C.byref(nf), C.byref(sg), C.byref(ns)))
        del keep
        return _take(fr, nf.value, FRAGMENT), _take(sg, ns.value, ALNSEG)

    def sort_adj_list_by_weighted_dfs(self, adj, start):
        """seq_db::sort_adj_list_by_weighted_dfs (seq_db.rs:1013-1061); start = (h0, h1, ori) -> DFSNODE[...]"""
        a = np.ascontiguousarray(adj, dtype=ADJ)
        st = np.zeros(1, dtype=GNODE)
        st["h0"], st["h1"], st["ori"] = start
        out, n = C.c_void_p(), C.c_size_t()
        _check(lib().pgr_b200_sort_adj_list_by_weighted_dfs(self.h, a.ctypes.data, a.size, st.ctypes.data, C.byref(out), C.byref(n)))
        return _take(out, n.value, DFSNODE)

    def get_principal_bundles_from_adj_list(self, adj, path_len_cutoff):
        """seq_db::get_principal_bundles_from_adj_list (seq_db.rs:1063-1186) -> (list of GNODE arrays, filtered ADJ[...])"""
        a = np.ascontiguousarray(adj, dtype=ADJ)
        v, off, flt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb, nf = C.c_size_t(), C.c_size_t()
        _check(lib().pgr_b200_principal_bundles(self.h, a.ctypes.data, a.size, path_len_cutoff, C.byref(v), C.byref(off), C.byref(nb),
                                                C.byref(flt), C.byref(nf)))
        off_a = _take(off, nb.value + 1, np.uint64)
        verts = _take(v, int(off_a[-1]), GNODE)
        return [verts[int(off_a[i]):int(off_a[i + 1])] for i in range(nb.value)], _take(flt, nf.value, ADJ)

    def get_principal_bundles(self, min_count, path_len_cutoff, keeps=None):
        """SeqIndexDB::get_principal_bundles (ext.rs:491-510): [] when the adjacency list is empty"""
        adj = self.adj_list(min_count, keeps)
        if adj.size == 0:
            return []
        return self.get_principal_bundles_from_adj_list(adj, path_len_cutoff)[0]


COMM_ID_BYTES = 128


def _take_query_result(res):
    """pgr_query_result* -> (q_target_off, target_sid, target_chain_off, chain_score, chain_hit_off, hits HITPAIR[...]); frees it"""
    r = res.contents

    def arr(ptr, n, dt):
        dt = np.dtype(dt)
        if n == 0:
            return np.zeros(0, dtype=dt)
        addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
        return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(addr), dtype=dt, count=n).copy()

    out = (arr(r.q_target_off, r.n_queries + 1, np.uint64), arr(r.target_sid, r.n_targets, np.uint32),
           arr(r.target_chain_off, r.n_targets + 1, np.uint64), arr(r.chain_score, r.n_chains, np.float32),
           arr(r.chain_hit_off, r.n_chains + 1, np.uint64), arr(r.hits, r.n_hits, HITPAIR))
    lib().pgr_b200_query_result_free(res)
    return out


class MdbMap:
    """the .mdb-resident look-up of the reference (read_mdb_file_to_frag_locations + raw_query_fragment_from_mmap_midx,
    seq_db.rs:1409-1471, :1230-1269): the key table in memory, the signatures left in the memory-mapped file"""

    def __init__(self, path):
        L = lib()
        L.pgr_b200_mdb_map_open.restype = C.c_void_p
        L.pgr_b200_mdb_map_open.argtypes = [C.c_char_p]
        L.pgr_b200_mdb_map_close.argtypes = [C.c_void_p]
        L.pgr_b200_mdb_map_info.argtypes = [C.c_void_p, C.POINTER(ShmmrSpec), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.pgr_b200_raw_query_mmap.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p),
                                              C.POINTER(C.c_void_p)]
        self.h = L.pgr_b200_mdb_map_open(str(path).encode())
        if not self.h:
            raise PgrError(-5, L.pgr_b200_last_error().decode())

    def info(self):
        spec, nk, ns = ShmmrSpec(), C.c_size_t(), C.c_size_t()
        _check(lib().pgr_b200_mdb_map_info(self.h, C.byref(spec), C.byref(nk), C.byref(ns)))
        return spec, nk.value, ns.value

    def raw_query(self, seq):
        """-> (pairs QPAIR[n], hit_off[n+1], hits SIG[...]), as ShmmrIndex.raw_query"""
        a = _bytes(seq)
        pairs, n, off, hits = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_void_p()
        _check(lib().pgr_b200_raw_query_mmap(self.h, a.ctypes.data, a.size, C.byref(pairs), C.byref(n), C.byref(off), C.byref(hits)))
        offs = _take(off, n.value + 1, np.uint64)
        return _take(pairs, n.value, QPAIR), offs, _take(hits, int(offs[-1]), SIG)

    def query_batch(self, seqs, penalty, max_count=None, max_count_query=None, max_count_target=None, max_aln_span=None,
                    max_gap=None, oriented=False, device=-1):
        """query_fragment_to_hps_from_mmap_file (ext.rs:285-342) for every sequence of `seqs`; result as ShmmrIndex.query_batch"""
        arrs, ptrs, lens = _seq_arrays(seqs)
        opt = lambda v: -1 if v is None else int(v)
        prm = QueryParams(penalty, opt(max_count), opt(max_count_query), opt(max_count_target), opt(max_aln_span), opt(max_gap), int(oriented))
        res = C.POINTER(QueryResult)()
        L = lib()
        L.pgr_b200_query_batch_mmap.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(QueryParams), C.POINTER(C.POINTER(QueryResult))]
        _check(L.pgr_b200_query_batch_mmap(self.h, device, len(arrs), ptrs, lens, C.byref(prm), C.byref(res)))
        return _take_query_result(res)

    def close(self):
        if self.h:
            lib().pgr_b200_mdb_map_close(self.h)
            self.h = None


def comm_unique_id():
    """NCCL unique id (bytes) made by one rank; broadcast it to the others by any means and pass it to Comm()"""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib().pgr_b200_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """one rank of the library's NCCL communicator (multi-GPU index build, one process per GPU)"""

    def __init__(self, unique_id, rank, n_ranks, device=-1):
        assert len(unique_id) == COMM_ID_BYTES
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self.h = lib().pgr_b200_comm_init_rank(buf, rank, n_ranks, device)
        if not self.h:
            _check(-3 if device_count() == 0 else -4)
        self.rank, self.n_ranks = rank, n_ranks

    def close(self):
        if getattr(self, "h", None):
            lib().pgr_b200_comm_free(self.h)
            self.h = None

    __del__ = close


class ShardedIndex:
    """ShmmrFragMap sharded by key range over the GPUs of this process (pgr_b200_mindex): one host thread and one NCCL
    rank per GPU, ONE all-to-all of the tuples at finalize.  devices=[0, 0] puts two shards on one GPU (test set-up)."""

    def __init__(self, spec, frg_id_mode=FRG_ID_FASTX, n_gpus=None, devices=None):
        if devices is not None:
            d = (C.c_int * len(devices))(*devices)
            self.h = lib().pgr_b200_mindex_new_devices(C.byref(spec), frg_id_mode, len(devices), d)
        else:
            self.h = lib().pgr_b200_mindex_new(C.byref(spec), frg_id_mode, n_gpus)
        if not self.h:
            _check(-3 if device_count() == 0 else -1)

    def close(self):
        if getattr(self, "h", None):
            lib().pgr_b200_mindex_free(self.h)
            self.h = None

    __del__ = close

    def n_shards(self):
        return lib().pgr_b200_mindex_n_shards(self.h)

    def add_batch(self, sids, seqs):
        arrs, ptrs, lens = _seq_arrays(seqs)
        s = np.ascontiguousarray(sids, dtype=np.uint32)
        _check(lib().pgr_b200_mindex_add_batch(self.h, len(arrs), s.ctypes.data, ptrs, lens))

    def finalize(self):
        _check(lib().pgr_b200_mindex_finalize(self.h))

    def counts(self):
        nk, ns, nf = C.c_size_t(), C.c_size_t(), C.c_uint32()
        _check(lib().pgr_b200_mindex_counts(self.h, C.byref(nk), C.byref(ns), C.byref(nf)))
        return nk.value, ns.value, nf.value

    def stats(self):
        self.finalize()
        out = []
        for g in range(self.n_shards()):
            st = ShardStats()
            _check(lib().pgr_b200_mindex_stats(self.h, g, C.byref(st)))
            out.append(st.as_dict())
        return out

    def shard(self, g):
        """shard g as a ShmmrIndex view (owned by this object: do not close it)"""
        h = lib().pgr_b200_mindex_shard(self.h, g)
        if not h:
            _check(-1)
        return _BorrowedIndex(h)

    def export(self):
        nk, ns, _ = self.counts()
        keys = np.zeros((max(nk, 1), 2), dtype=np.uint64)
        offs = np.zeros(nk + 1, dtype=np.uint64)
        sigs = np.zeros(max(ns, 1), dtype=SIG)
        _check(lib().pgr_b200_mindex_export_csr(self.h, keys.ctypes.data, offs.ctypes.data, sigs.ctypes.data))
        return keys[:nk], offs, sigs[:ns]

    def write_mdb(self, path):
        _check(lib().pgr_b200_mindex_write_mdb(self.h, path.encode()))

    def gather(self, device=0):
        """the whole map as one ordinary ShmmrIndex on `device` (for queries, adjacency, fragment compression)"""
        h = lib().pgr_b200_mindex_gather(self.h, device)
        if not h:
            _check(-4)
        return ShmmrIndex(handle=h)


class _BorrowedIndex(ShmmrIndex):
    """a ShmmrIndex whose handle belongs to a ShardedIndex"""

    def __init__(self, h):
        self.h = h

    def close(self):
        self.h = None

    __del__ = close


def sparse_aln(hits, max_span, penalty, max_gap=None, oriented=False):
    """aln::sparse_aln (aln.rs:12-142) -> (scores, chain_off, chain_hits, sorted_hits)"""
    h = np.ascontiguousarray(hits, dtype=HITPAIR).copy()
    nc = C.c_size_t()
    off, sc, ch = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _check(lib().pgr_b200_sparse_aln(h.ctypes.data, h.size, max_span, penalty, -1 if max_gap is None else int(max_gap), int(oriented),
                                     C.byref(nc), C.byref(off), C.byref(sc), C.byref(ch)))
    off_a = _take(off, nc.value + 1, np.uint64)
    return _take(sc, nc.value, np.float32), off_a, _take(ch, int(off_a[-1]), HITPAIR), h
