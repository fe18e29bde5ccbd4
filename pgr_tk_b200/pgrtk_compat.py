"""pgrtk_compat — the Python surface of the reference (pgr-tk/src/lib.rs, the `pgrtk` extension module) re-hosted on the
C ABI through ctypes: same class / method / function names, argument order and return shapes, for the part of the API
that lies on the SHIMMER indexing path.  Every method cites the reference method it mirrors; anything outside the path
(AGC back end, WFA, consensus, ...) is absent and raises AttributeError like any missing attribute.

Where the reference returns hash-map iteration order the canonical orders of DESIGN.md are used (targets by sid, keys
ascending).  FASTX back end only: sequences are held in host memory, the index on the GPU.
"""
import gzip

import numpy as np

from . import api


def _read_fastx(path):
    """fasta_io.rs:46-172 (FASTA part): id = header up to the first space, sequence = everything up to the next '>'"""
    opener = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    data = opener(path, "rb").read()
    recs = []
    if not data:
        return recs
    for chunk in data[1:].split(b">"):
        head, _, body = chunk.partition(b"\n")
        name = head.split(b" ")[0].replace(b"\r", b"").decode()
        recs.append((name, body.replace(b"\n", b"").replace(b"\r", b"")))
    return recs


class SeqIndexDB:
    """pgr-tk/src/lib.rs:58-1425 `SeqIndexDB`"""

    def __init__(self):
        self._idx = None
        self._spec = None
        self._seqs = []          # (ctg_name, source, bytes) by sid
        self.seq_index = None    # (ctg_name, source) -> (sid, len)     lib.rs:215
        self.seq_info = None     # sid -> (ctg_name, source, len)       lib.rs:223

    # ---- loading (lib.rs:142-213) ----------------------------------------------------------------------------------------
    def load_from_fastx(self, filepath, w=80, k=56, r=4, min_span=64):
        self._spec = api.ShmmrSpec(w, k, r, min_span)
        self._idx = api.ShmmrIndex(self._spec, api.FRG_ID_FASTX)
        self._seqs, self.seq_index, self.seq_info = [], {}, {}
        self._add(_read_fastx(filepath), filepath)

    def append_from_fastx(self, filepath):
        if self._idx is None:
            raise RuntimeError("Only DB created with load_from_fastx() can add data from another fastx file")
        self._add(_read_fastx(filepath), filepath)

    def load_from_seq_list(self, seq_list, source="Memory", w=80, k=56, r=4, min_span=8):
        self._spec = api.ShmmrSpec(w, k, r, min_span)
        self._idx = api.ShmmrIndex(self._spec, api.FRG_ID_FASTX)
        self._seqs, self.seq_index, self.seq_info = [], {}, {}
        self._add([(n, bytes(s)) for n, s in seq_list], source)

    def _add(self, recs, source):
        sid0 = len(self._seqs)
        for i, (name, seq) in enumerate(recs):
            self._seqs.append((name, source, seq))
            self.seq_index[(name, source)] = (sid0 + i, len(seq))
            self.seq_info[sid0 + i] = (name, source, len(seq))
        if recs:
            self._idx.add_batch(list(range(sid0, sid0 + len(recs))), [s for _, s in recs])

    # ---- queries ------------------------------------------------------------------------------------------------------------
    def query_fragment(self, seq):
        """lib.rs:249-305 -> [((h0, h1), (bgn, end, orientation), [FragmentSignature, ...]), ...]"""
        pairs, off, hits = self._idx.raw_query(bytes(seq))
        out = []
        for i, p in enumerate(pairs):
            sigs = [(int(h["frg_id"]), int(h["sid"]), int(h["bgn"]), int(h["end"]), int(h["ori"])) for h in hits[int(off[i]):int(off[i + 1])]]
            # raw_query_fragment returns EVERY query pair, with an empty vector when the key is absent (seq_db.rs:1219-1226)
            out.append(((int(p["h0"]), int(p["h1"])), (int(p["bgn"]), int(p["end"]), int(p["ori"])), sigs))
        return out

    def query_fragment_to_hps(self, seq, penalty, max_count=None, max_count_query=None, max_count_target=None, max_aln_span=None,
                              max_gap=None, orientated=False):
        """lib.rs:365-420 -> [(sid, [(score, [((q_bgn, q_end, q_ori), (t_bgn, t_end, t_ori)), ...]), ...]), ...]"""
        tsid, tco, csc, cho, hits = self._idx.query_fragment_to_hps(bytes(seq), penalty, max_count=max_count, max_count_query=max_count_query,
                                                                    max_count_target=max_count_target, max_aln_span=max_aln_span,
                                                                    max_gap=max_gap, oriented=bool(orientated))
        out = []
        for t, sid in enumerate(tsid):
            alns = []
            for c in range(int(tco[t]), int(tco[t + 1])):
                hp = [((int(h["qb"]), int(h["qe"]), int(h["qo"])), (int(h["tb"]), int(h["te"]), int(h["to"]))) for h in hits[int(cho[c]):int(cho[c + 1])]]
                alns.append((float(csc[c]), hp))
            out.append((int(sid), alns))
        return out

    # ---- the index ------------------------------------------------------------------------------------------------------------
    def get_shmmr_spec(self):
        """lib.rs:729-750"""
        s = self._spec
        return None if s is None else (s.w, s.k, s.r, s.min_span, bool(s.sketch))

    def get_shmmr_map(self):
        """lib.rs:752-772 -> {(h0, h1): [(frg_id, sid, bgn, end, orientation), ...]}"""
        return self._idx.as_map()

    def get_shmmr_pair_count(self, shmmr_pair):
        """lib.rs:636-666"""
        return len(self._idx.as_map().get((int(shmmr_pair[0]), int(shmmr_pair[1])), []))

    def get_shmmr_pair_list(self):
        """lib.rs:774-790 -> [(h0, h1, sid, bgn, end, orientation), ...] (keys ascending)"""
        keys, offs, sigs = self._idx.export()
        rep = np.repeat(np.arange(len(keys)), np.diff(offs.astype(np.int64)))
        return [(int(keys[k, 0]), int(keys[k, 1]), int(s["sid"]), int(s["bgn"]), int(s["end"]), int(s["ori"])) for k, s in zip(rep, sigs)]

    # ---- sequences (lib.rs:810-891) ------------------------------------------------------------------------------------------
    def get_seq_by_id(self, sid):
        return list(self._seqs[sid][2])

    def get_sub_seq_by_id(self, sid, bgn, end):
        return list(self._seqs[sid][2][bgn:end])

    def get_seq(self, sample_name, ctg_name):
        return self.get_seq_by_id(self.seq_index[(ctg_name, sample_name)][0])

    def get_sub_seq(self, sample_name, ctg_name, bgn, end):
        return self.get_sub_seq_by_id(self.seq_index[(ctg_name, sample_name)][0], bgn, end)

    # ---- MAP graph (lib.rs:893-1064) -----------------------------------------------------------------------------------------
    def get_smp_adj_list(self, min_count, keeps=None):
        """-> [(sid, (h0, h1, ori), (h0, h1, ori)), ...]"""
        return [(int(a["sid"]), (int(a["a0"]), int(a["a1"]), int(a["ori0"])), (int(a["b0"]), int(a["b1"]), int(a["ori1"])))
                for a in self._idx.adj_list(min_count, keeps)]

    def get_smp_adj_list_by_seq(self, min_count, keeps=None):
        """the adjacency list generate_mapg_gfa builds when method != "from_fragmap" (ext.rs:696-722): every sequence's own
        shimmer pairs through generate_smp_adj_list_for_seq (seq_db.rs:946-1000), min_count 0 for the sequences in `keeps`;
        sequences in ascending sid order (the reference walks seq_info's hash map: order unpinned)"""
        sids = sorted(self.seq_info)
        ks = set(keeps) if keeps is not None else set()
        mcs = [0 if sid in ks else min_count for sid in sids]
        adj = self._idx.smp_adj_list_for_seqs(sids, [self._seqs[sid][2] for sid in sids], mcs)
        return [(int(a["sid"]), (int(a["a0"]), int(a["a1"]), int(a["ori0"])), (int(a["b0"]), int(a["b1"]), int(a["ori1"]))) for a in adj]

    def generate_mapg_gfa(self, min_count, filepath, method="from_fragmap", keeps=None):
        """lib.rs:1066-1080 -> ext.rs:652-786.  S lines by segment id, L lines by first appearance in the adjacency list (the
        reference iterates FxHashMaps: order unpinned)"""
        adj = self.get_smp_adj_list(min_count, keeps) if method == "from_fragmap" else self.get_smp_adj_list_by_seq(min_count, keeps)
        fmap = self.get_shmmr_map()
        k = self.get_shmmr_spec()[1]
        overlaps, frag_id = {}, {}
        for sid, v, w in adj:
            if v[0] <= w[0]:                                   # ext.rs:724
                overlaps.setdefault((v, w), []).append((sid, v[2], w[2]))
                frag_id.setdefault((v[0], v[1]), len(frag_id))
                frag_id.setdefault((w[0], w[1]), len(frag_id))
        with open(filepath, "w") as f:
            f.write("H\tVN:Z:1.0\tCM:Z:Sparse Genome Graph Generated By pgr-tk\n")
            for smp, i in sorted(frag_id.items(), key=lambda t: t[1]):
                hits = fmap[smp]
                ave_len = (sum(h[3] - h[2] for h in hits) & 0xFFFFFFFF) // len(hits)
                f.write("S\t%d\t*\tLN:i:%d\tSN:Z:%016x_%016x\n" % (i, ave_len + k, smp[0], smp[1]))
            for (v, w), vs in overlaps.items():
                f.write("L\t%d\t%s\t%d\t%s\t%dM\tSC:i:%d\n" % (frag_id[(v[0], v[1])], "+-"[v[2]], frag_id[(w[0], w[1])], "+-"[w[2]], k, len(vs)))

    def _adj_array(self, adj_list):
        a = np.zeros(len(adj_list), dtype=api.ADJ)
        for i, (sid, v, w) in enumerate(adj_list):
            a[i]["sid"], a[i]["a0"], a[i]["a1"], a[i]["ori0"], a[i]["b0"], a[i]["b1"], a[i]["ori1"] = sid, v[0], v[1], v[2], w[0], w[1], w[2]
        return a

    def sort_adj_list_by_weighted_dfs(self, adj_list, start):
        """lib.rs:938-1000 -> [(node, previous node or None, weight, is_leaf, rank, branch, branch_rank), ...]"""
        rows = self._idx.sort_adj_list_by_weighted_dfs(self._adj_array(adj_list), start)
        g = lambda v: (int(v["h0"]), int(v["h1"]), int(v["ori"]))
        return [(g(r["node"]), g(r["prev"]) if r["has_prev"] else None, int(r["weight"]), bool(r["is_leaf"]), int(r["rank"]), int(r["branch"]),
                 int(r["branch_rank"])) for r in rows]

    def get_principal_bundles(self, min_count, path_len_cutoff, keeps=None):
        """lib.rs:1003-1064 (ext.rs:491-510) -> [[(h0, h1, ori), ...], ...]"""
        return [[(int(v["h0"]), int(v["h1"]), int(v["ori"])) for v in b] for b in self._idx.get_principal_bundles(min_count, path_len_cutoff, keeps)]


def sparse_aln(sp_hits, max_span, penalty, max_gap=None, orientated=False):
    """lib.rs:1538-1579 -> [(score, [HitPair, ...]), ...]"""
    h = np.zeros(len(sp_hits), dtype=api.HITPAIR)
    for i, (q, t) in enumerate(sp_hits):
        h[i]["qb"], h[i]["qe"], h[i]["qo"], h[i]["tb"], h[i]["te"], h[i]["to"] = q[0], q[1], q[2], t[0], t[1], t[2]
    sc, off, ch, _ = api.sparse_aln(h, max_span, penalty, max_gap, orientated)
    return [(float(sc[c]), [((int(x["qb"]), int(x["qe"]), int(x["qo"])), (int(x["tb"]), int(x["te"]), int(x["to"]))) for x in ch[int(off[c]):int(off[c + 1])]])
            for c in range(len(sc))]


def get_shmmr_pairs_from_seq(seq, w=80, k=56, r=4, min_span=16, padding=False):
    """lib.rs:1581-1613 -> [(h0, h1, bgn, end, orientation), ...] (strict '<' canonical form)"""
    mm = api.sequence_to_shmmrs(0, bytes(seq), api.ShmmrSpec(w, k, r, min_span), padding)
    out = []
    for a, b in zip(mm[:-1], mm[1:]):
        s0, s1 = int(a["x"]) >> 8, int(b["x"]) >> 8
        p0, p1 = ((int(a["y"]) & 0xFFFFFFFF) >> 1) + 1, ((int(b["y"]) & 0xFFFFFFFF) >> 1) + 1
        out.append((s0, s1, p0, p1, 0) if s0 < s1 else (s1, s0, p0, p1, 1))
    return out
