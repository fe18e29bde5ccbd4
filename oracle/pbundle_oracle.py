"""pbundle_oracle.py — TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of the principal-bundle bookkeeping of
pgr-pbundle-decomp: ext.rs:552-650 (get_principal_bundles_with_id), ext.rs:976-1015 (get_principal_bundle_decomposition),
pgr-bin/src/bin/pgr-pbundle-decomp.rs:62-137 (group_smps_by_principle_bundle_id) and :337-529 (the .bed and
.ctg.summary.tsv writers).  f32 arithmetic goes through numpy.float32; Rust's `{}` of an f32 is the shortest
round-trip decimal without exponent.  PARITY UNPINNED (no reference fixture); the bundles themselves come from
bundles_oracle.py.  A smp is (h0, h1, bgn, end, ori); a bundle vertex (h0, h1, ori).
"""
import numpy as np

f32 = np.float32


def f32_display(v):
    v = f32(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def principal_bundles_with_id(pb, smps_by_sid):
    """ext.rs:552-650 -> ([(bundle_id, mean_order, vertices)], vertex_map {(h0,h1): (bid, ori, pos)})"""
    vmap = {}
    for bid, path in enumerate(pb):
        for p, v in enumerate(path):
            vmap[(v[0], v[1])] = (bid, v[2], p)
    directions, orders = {}, {}
    for sid in sorted(smps_by_sid):
        visited = set()
        for order, v in enumerate(smps_by_sid[sid]):
            b = vmap.get((v[0], v[1]))
            if b is None:
                continue
            if b[0] not in visited:
                orders.setdefault(b[0], []).append(f32(order))
                visited.add(b[0])
            directions.setdefault(b[0], []).append(0 if b[1] == v[4] else 1)
    mod = []
    for bid in range(len(pb)):
        if bid in orders:
            s = f32(0.0)
            for o in orders[bid]:
                s = f32(s + o)
            mean = int(f32(s / f32(len(orders[bid]))))
            d = directions[bid]
            mod.append((mean, bid, 0 if sum(d) < (len(d) >> 1) else 1))
        else:
            mod.append((2 ** 64 - 1, bid, 0))
    mod.sort()
    out = []
    for ord_, bid, direction in mod:
        if direction == 1:
            rpb = [(v[0], v[1], 1 - v[2]) for v in reversed(pb[bid])]
            for p, v in enumerate(rpb):
                vmap[(v[0], v[1])] = (bid, v[2], p)
            out.append((bid, ord_, rpb))
        else:
            out.append((bid, ord_, list(pb[bid])))
    return out, vmap


def group_smps(smps, vmap, cutoff, merge_distance):
    """pgr-pbundle-decomp.rs:62-137 -> [[(smp, bid, d, bpos), ...], ...]"""
    pre = None
    allp, cur = [], []
    for smp in smps:
        info = vmap.get((smp[0], smp[1]))
        if info is None:
            continue
        d = 0 if smp[4] == info[1] else 1
        bid, bpos = info[0], info[2]
        if pre is None:
            cur = [(smp, bid, d, bpos)]
            pre = (bid, d)
            continue
        if (bid, d) != pre:
            if cur[-1][0][3] - cur[0][0][2] > cutoff:
                allp.append(cur)
            cur = []
            pre = (bid, d)
        cur.append((smp, bid, d, bpos))
    if cur and cur[-1][0][3] - cur[0][0][2] > cutoff:
        allp.append(cur)
    if not allp:
        return []
    rtn = []
    part = list(allp[0])
    for p in allp[1:]:
        last = part[-1]
        if last[1] == p[0][1] and last[2] == p[0][2] and abs(p[0][0][2] - last[0][3]) < merge_distance:
            part.extend(p)
        else:
            rtn.append(part)
            part = list(p)
    if part:
        rtn.append(part)
    return rtn


def decomposition_files(cmd_string, seq_info, smps_by_sid, pbid, vmap, k, cutoff, merge_distance):
    """seq_info = [(sid, len, ctg)]; -> (bed text, summary text)"""
    bid_to_size = {b[0]: len(b[2]) for b in pbid}
    order = sorted(seq_info, key=lambda t: t[2])
    bed = ["# cmd: %s" % cmd_string]
    rep, non = {}, {}
    for sid, _ln, ctg in order:
        parts = group_smps(smps_by_sid[sid], vmap, cutoff, merge_distance)
        cnt = {}
        for p in parts:
            cnt[p[0][1]] = cnt.get(p[0][1], 0) + 1
        for p in parts:
            b, e = p[0][0][2], p[-1][0][3] + k
            bid = p[0][1]
            r = cnt[bid] > 1
            (rep if r else non).setdefault(sid, []).append(e - b - k)
            bed.append("%s\t%d\t%d\t%d:%d:%d:%d:%d:%s" % (ctg, b, e, bid, bid_to_size[bid], p[0][2], p[0][3], p[-1][3], "R" if r else "U"))
    hdr = ["ctg", "length", "repeat_bundle_count", "repeat_bundle_sum", "repeat_bundle_percentage", "repeat_bundle_mean", "repeat_bundle_min",
           "repeat_bundle_max", "non_repeat_bundle_count", "non_repeat_bundle_sum", "non_repeat_bundle_percentage", "non_repeat_bundle_mean",
           "non_repeat_bundle_min", "non_repeat_bundle_max", "total_bundle_count", "total_bundle_coverage_percentage"]
    summ = ["#" + "\t".join(hdr)]

    def stats(v):
        if not v:
            return 0, "NA", "NA", "NA"
        s = sum(v)
        return s, f32_display(f32(s) / f32(len(v))), str(min(v)), str(max(v))

    def pct(x, ln):
        return f32_display(f32(f32(100.0) * f32(x)) / f32(ln))

    with np.errstate(divide="ignore", invalid="ignore"):
        for sid, ln, ctg in order:
            rv, nv = rep.get(sid, []), non.get(sid, [])
            rs, rmean, rmin, rmax = stats(rv)
            ns, nmean, nmin, nmax = stats(nv)
            summ.append("\t".join(map(str, [ctg, ln, len(rv), rs, pct(rs, ln), rmean, rmin, rmax, len(nv), ns, pct(ns, ln), nmean, nmin, nmax,
                                            len(rv) + len(nv), pct(rs + ns, ln)])))
    return "\n".join(bed) + "\n", "\n".join(summ) + "\n"


# ---- MAP-graph text files (ext.rs:652-959); canonical line orders where the reference walks FxHashMaps: S lines by segment
# id, L lines by first appearance in the adjacency list, C lines by sequence id, F lines by key ascending ----------------
def gfa_text(adj_list, frag_map, k, vmap=None):
    """adj_list = [(sid, v, w)], v = (h0, h1, ori); frag_map[(h0, h1)] = [(frg_id, sid, bgn, end, ori)] -> GFA text
    (generate_mapg_gfa ext.rs:735-786; with vmap = {(h0, h1): (bundle, ori, pos)} generate_principal_mapg_gfa :885-956)"""
    overlaps, frag_id = {}, {}
    for sid, v, w in adj_list:
        if v[0] <= w[0]:
            overlaps.setdefault((v, w), []).append((sid, v[2], w[2]))
            frag_id.setdefault((v[0], v[1]), len(frag_id))
            frag_id.setdefault((w[0], w[1]), len(frag_id))
    out = ["H\tVN:Z:1.0\tCM:Z:Sparse Genome Graph Generated By pgr-tk"]
    for smp, i in sorted(frag_id.items(), key=lambda t: t[1]):
        hits = frag_map[smp]
        ave_len = (sum(h[3] - h[2] for h in hits) & 0xFFFFFFFF) // len(hits)
        line = "S\t%d\t*\tLN:i:%d\tSN:Z:%016x_%016x" % (i, ave_len + k, smp[0], smp[1])
        if vmap is not None and smp in vmap:
            line += "\tBN:i:%d\tBP:i:%d" % (vmap[smp][0], vmap[smp][2])
        out.append(line)
    for (v, w), vs in overlaps.items():
        out.append("L\t%d\t%s\t%d\t%s\t%dM\tSC:i:%d" % (frag_id[(v[0], v[1])], "+-"[v[2]], frag_id[(w[0], w[1])], "+-"[w[2]], k, len(vs)))
    return "\n".join(out) + "\n"


def mapg_idx_text(spec, seq_info, frag_map):
    """ext.rs:791-847; spec = (w, k, r, min_span, sketch); seq_info = [(sid, len, ctg, source)]"""
    out = ["K\t%d\t%d\t%d\t%d\t%s" % (spec[0], spec[1], spec[2], spec[3], "true" if spec[4] else "false")]
    for sid, ln, ctg, src in sorted(seq_info):
        out.append("C\t%d\t%s\t%s\t%d" % (sid, ctg, src if src else "NA", ln))
    for key in sorted(frag_map):
        for v in frag_map[key]:
            out.append("F\t%016x_%016x\t%d\t%d\t%d\t%d\t%d" % (key[0], key[1], v[0], v[1], v[2], v[3], v[4]))
    return "\n".join(out) + "\n"


# ---- the .pdb file (pgr-pbundle-decomp.rs:158-226, :367-396): "PDB:0.5" + bincode 2 standard config of
# (w, k, r, min_span, min_branch_size, min_cov, [(bundle_id, mean_order, [(h0, h1, ori)])], {(h0, h1): (bundle, ori, pos)}) ----
def _vi(v):
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + v.to_bytes(2, "little")
    if v < 1 << 32:
        return b"\xfc" + v.to_bytes(4, "little")
    return b"\xfd" + v.to_bytes(8, "little")


def encode_pdb(w, k, r, min_span, min_branch_size, min_cov, pbid, vmap):
    o = [b"PDB:0.5", _vi(w), _vi(k), _vi(r), _vi(min_span), _vi(min_branch_size), _vi(min_cov), _vi(len(pbid))]
    for bid, order, verts in pbid:
        o += [_vi(bid), _vi(order), _vi(len(verts))]
        for v in verts:
            o += [_vi(v[0]), _vi(v[1]), bytes([v[2]])]
    o.append(_vi(len(vmap)))
    for key in sorted(vmap):                      # the reference writes FxHashMap order; canonical: ascending key
        b = vmap[key]
        o += [_vi(key[0]), _vi(key[1]), _vi(b[0]), bytes([b[1]]), _vi(b[2])]
    return b"".join(o)


def decode_pdb(buf):
    assert buf[:7] == b"PDB:0.5"
    pos = [7]

    def u8():
        pos[0] += 1
        return buf[pos[0] - 1]

    def vi():
        t = u8()
        if t < 251:
            return t
        n = {251: 2, 252: 4, 253: 8, 254: 16}[t]
        v = int.from_bytes(buf[pos[0]:pos[0] + n], "little")
        pos[0] += n
        return v
    head = tuple(vi() for _ in range(6))
    pbid = []
    for _ in range(vi()):
        bid, order = vi(), vi()
        pbid.append((bid, order, [(vi(), vi(), u8()) for _ in range(vi())]))
    vmap = {}
    for _ in range(vi()):
        key = (vi(), vi())
        vmap[key] = (vi(), u8(), vi())
    assert pos[0] == len(buf)
    return head, pbid, vmap
