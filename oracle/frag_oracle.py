"""frag_oracle.py — TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of the reference's fragment compression:
  * shmmrutils.rs:35-54   track_delta_point
  * shmmrutils.rs:57-223  match_reads (the O(nD) variant with banding)
  * seq_db.rs:113-156     deltas_to_aln_segs
  * seq_db.rs:189-358     CompactSeqDB::seq_to_compressed (prefix / internal / suffix fragments, alignment of an internal
                          fragment against the first earlier `Internal` fragment of its shimmer pair that matches)
  * seq_db.rs:507-525     load_seqs_from_seq_vec (sequences in order; the frag_map a sequence sees holds earlier sequences only)
Pinned by tests/test_frag_format.py: on test_seqs.fa it reproduces, fragment by fragment, the Vec<Fragment> stored in the
reference's committed test_seqs_frag.frg (and therefore its .sdx chunk table and the inflated chunk payloads byte for byte).
Shimmers come from the C++ oracle (tests/orc.py).  Small inputs only (pure-Python loops).
"""
from frag_format import FRAG_ALN, FRAG_INTERNAL, FRAG_PREFIX, FRAG_SUFFIX, SEG_FULL, SEG_INS, SEG_MATCH, reverse_complement


def match_reads(seq0, seq1, get_delta, tol, min_match_len, min_match_start, bandwidth):
    """shmmrutils.rs:57-223 -> None or dict(bgn0, end0, bgn1, end1, m_end0, m_end1, dist, m_size, deltas=[(x, y, dk)])"""
    len0, len1 = len(seq0), len(seq1)
    d_max = 32 + int(tol * (len0 if len0 < len1 else len1))
    k_min = k_max = 0
    uv = {d: (0, 0) for d in range(-d_max, d_max + 1)}
    delta_pts = {}
    best_m = -1
    matched = False
    d_final = k_final = 0
    start = False
    longest = 0
    rtn = dict(m_size=0, dist=0, bgn0=0, end0=0, bgn1=0, end1=0, m_end0=0, m_end1=0, deltas=None)
    for d in range(d_max):
        if k_max - k_min > bandwidth:
            break
        for k in range(k_min, k_max + 1, 2):
            vn = uv[k - 1][1]
            vp = uv[k + 1][1]
            if k == k_min or (k != k_max and vn < vp):
                x, pre_k = vp, k + 1
            else:
                x, pre_k = vn + 1, k - 1
            y = (x - k) & 0xFFFFFFFF
            if get_delta:
                delta_pts.setdefault((d, k), (x, y, k - pre_k))
            x1, y1 = x, y
            while x < len0 and y < len1 and seq0[x] == seq1[y]:
                x += 1
                y += 1
            if (x - x1) >= min_match_start and not start:
                rtn["bgn0"], rtn["bgn1"] = x1, y1
                start = True
            if (x - x1) > longest:
                longest = x - x1
                rtn["m_end0"], rtn["m_end1"] = x, y
            uv[k] = ((x + y) & 0xFFFFFFFF, x)
            if x + y > best_m:
                best_m = x + y
            if x >= len0 or y >= len1:
                matched = True
                d_final, k_final = d, k
                rtn["end0"], rtn["end1"] = x, y
                break
        k_max_new, k_min_new = k_min, k_max
        for k2 in range(k_min, k_max + 1, 2):
            if uv[k2][0] >= best_m - bandwidth:
                k_min_new = min(k_min_new, k2)
                k_max_new = max(k_max_new, k2)
        k_max, k_min = k_max_new + 1, k_min_new - 1
        if matched:
            d_inside = 0
            if get_delta:
                dpts = []
                dd, kk = d_final, k_final
                while dd > 0:                         # track_delta_point
                    p = delta_pts[(dd, kk)]
                    if rtn["bgn0"] <= p[0] <= rtn["end0"]:
                        dpts.append(p)
                    dd -= 1
                    kk -= p[2]
                d_inside = sum(1 for p in dpts if rtn["bgn0"] < p[0] < rtn["end0"])
                rtn["deltas"] = dpts
            rtn["dist"] = d_inside
            rtn["m_size"] = (rtn["end0"] - rtn["bgn0"] + rtn["end1"] - rtn["bgn1"] + 2 * d_inside) >> 1
            if rtn["m_size"] < min_match_len:
                matched = False
            break
    return rtn if matched else None


def deltas_to_aln_segs(deltas, endx, endy, base_frg, frg):
    """seq_db.rs:113-156"""
    if not deltas and len(base_frg) == len(frg):
        return [(SEG_FULL,)]
    segs = []
    x, y = endx, endy
    for yy in range(len(frg) - 1, y - 1, -1):
        segs.append((SEG_INS, frg[yy]))
    for x1, y1, dk in deltas:
        if x1 < x:
            segs.append((SEG_MATCH, x1, x))
        x, y = x1, y1
        if dk > 0:
            x -= dk
        else:
            for yy in range(-dk):
                segs.append((SEG_INS, frg[y - yy - 1]))
    if x != 0:
        segs.append((SEG_MATCH, 0, x))
    segs.reverse()
    return segs


class CompactSeqDB:
    """the fragment side of seq_db.rs CompactSeqDB (FASTX path)"""

    def __init__(self, k):
        self.k = k
        self.frags = []
        self.frag_map = {}      # (h0, h1) -> [(frg_id, sid, bgn, end, orientation)]
        self.seqs = []

    def seq_to_compressed(self, source, name, sid, seq, shmmrs, try_compress=True):
        """shmmrs = [(hash, pos)] of the sequence (MM128.hash() = x >> 8, pos())"""
        frags, k = self.frags, self.k
        frg_id = len(frags)
        first = frg_id
        if not shmmrs:
            frags.append((FRAG_PREFIX, bytes(seq)))
            frags.append((FRAG_SUFFIX, b""))
            self.seqs.append(dict(source=source, name=name, id=sid, seq_frag_range=(first, 2), len=len(seq)))
            return
        n_frags = 1
        frags.append((FRAG_PREFIX, bytes(seq[:shmmrs[0][1] + 1])))
        frg_id += 1
        internal = []
        for (s0, p0), (s1, p1) in zip(shmmrs[:-1], shmmrs[1:]):
            pair, orientation = ((s0, s1), 0) if s0 <= s1 else ((s1, s0), 1)
            bgn, end = p0 + 1, p1 + 1
            out = None
            if end - bgn > 128 and try_compress and pair in self.frag_map:
                for t in self.frag_map[pair]:
                    base = frags[t[0]]
                    if base[0] != FRAG_INTERNAL:
                        continue
                    frg = bytes(seq[bgn - k:end])
                    rc = orientation != t[4]
                    if rc:
                        frg = reverse_complement(frg)
                    m = match_reads(base[1], frg, True, 0.1, 0, 0, 32)
                    if m is None:
                        continue
                    segs = deltas_to_aln_segs(m["deltas"], m["end0"], m["end1"], base[1], frg)
                    if 8 > (len(frg) >> 2):               # align_of_val(&Vec) = 8 (seq_db.rs:306)
                        continue
                    out = (pair, (FRAG_ALN, t[0], rc, len(frg), segs), bgn, end, orientation)
                    break
            if out is None:
                out = (pair, (FRAG_INTERNAL, bytes(seq[bgn - k:end])), bgn, end, orientation)
            internal.append(out)
        for pair, frg, bgn, end, orientation in internal:
            self.frag_map.setdefault(pair, []).append((frg_id, sid, bgn, end, orientation))
            frags.append(frg)
            frg_id += 1
            n_frags += 1
        frags.append((FRAG_SUFFIX, bytes(seq[shmmrs[-1][1] + 1:])))
        n_frags += 1
        self.seqs.append(dict(source=source, name=name, id=sid, seq_frag_range=(first, n_frags), len=len(seq)))
