"""pgr_tk_b200 — host-side mirror (ctypes over the C ABI in include/pgr_b200.h) of the pgr-db interfaces on the
SHIMMER indexing hot path: shmmrutils::sequence_to_shmmrs (shmmrutils.rs:657), CompactSeqDB (seq_db.rs:95),
raw_query_fragment (seq_db.rs:1200), aln::query_fragment_to_hps (aln.rs:147), frag_map_to_adj_list (seq_db.rs:876).

There is no CPU fallback: importing works anywhere (so that the CPU test-suite can check the ABI), but every compute
call raises PgrError when libpgr_b200.so is missing or no B200 is visible.
"""
from .api import (  # noqa: F401
    FRG_ID_AGC, FRG_ID_FASTX, TUPLE, Comm, MdbMap, ShardedIndex, ShardStats, comm_unique_id, QueryParams, ShmmrIndex, sparse_aln, ADJ, ALNSEG, FRAGMENT, GNODE, DFSNODE, HITPAIR, MM128, QPAIR, SIG, Ctx, PgrError, ShmmrSpec, build_library, device_count, get_shmmrs_from_seqs,
    host_alloc, lib, library_path, pack_bases, pack_isa, set_transport, TRANSPORT_PACKED, TRANSPORT_DIRECT, sequence_to_shmmrs, set_default_device,
)
