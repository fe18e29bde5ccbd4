"""query_post_oracle.py — TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of pgr-query's post-processing
(pgr-bin/src/bin/pgr-query.rs:166-430): chains with more than two anchors -> per-target ranges -> merged forward /
reverse regions -> the lines of <out>.NNN.hit / .hit.bed and the records of <out>.NNN.fa.

The reference iterates FxHashMaps over target ids (pgr-query.rs:167,187,200,348); no test pins that order, so targets
are taken in ascending sid order here and in the product (PARITY UNPINNED for the order of targets only).
A HitPair is ((q_bgn, q_end, q_ori), (t_bgn, t_end, t_ori)).
"""
import os


def merge_regions(rgns, tol):
    """pgr-query.rs:219-245 / 251-277; rgns sorted tuples (bgn, end, len, orientation, aln)"""
    out = []
    last = (0, 0, 0, 0, [])
    for r in rgns:
        if not last[4]:
            last = r
            continue
        l_bgn, l_end = last[0], last[1]
        assert l_end > l_bgn
        r_bgn, r_end = r[0], r[1]
        if r_bgn - l_end < tol:
            end = r_end if r_end > l_end else l_end
            last = (l_bgn, end, end - l_bgn, last[3], last[4] + r[4])
        else:
            out.append(last)
            last = r
    if last[2] > 0:
        out.append(last)
    return out


def merge_query_hits(targets, tol):
    """targets: [(sid, [(score, [HitPair, ...]), ...])] as query_fragment_to_hps returns them -> [(sid, [region, ...])]"""
    out = []
    for sid, alns in sorted(targets, key=lambda t: t[0]):
        f_count = r_count = 0                     # declared per target, never reset (pgr-query.rs:170-171)
        rgns = []
        for _score, aln in alns:
            if len(aln) > 2:
                for hp in aln:
                    if hp[0][2] == hp[1][2]:
                        f_count += 1
                    else:
                        r_count += 1
                orientation = 0 if f_count > r_count else 1
                tc = sorted((hp[1][0], hp[1][1]) for hp in aln)
                bgn, end = tc[0][0], tc[-1][1]
                rgns.append((bgn, end, end - bgn, orientation, list(aln)))
        if not rgns:
            continue
        f_rgns = sorted(r for r in rgns if r[3] == 0)
        r_rgns = sorted(r for r in rgns if r[3] == 1)
        out.append((sid, merge_regions(f_rgns, tol) + merge_regions(r_rgns, tol)))
    return out


def file_stem(path):
    name = os.path.basename(path)
    dot = name.rfind(".")
    return name if dot <= 0 else name[:dot]


COMP = {ord(a): ord(b) for a, b in zip("ACGTacgt", "TGCAtgca")}


def reverse_complement(seq):
    return bytes(COMP.get(b, b) for b in reversed(seq))


def hit_lines(idx, q_name, q_len, merged, seq_info, bed=False):
    """-> (text of the hit file, [(target_seq_name, sid, bgn, end, orientation)])"""
    if bed:
        lines = ["#" + "\t".join(["target", "bgn", "end", "query", "color", "orientation", "q_len", "aln_anchor_count", "q_idx", "src", "ctg_bgn", "ctg_end"])]
    else:
        lines = ["#" + "\t".join(["idx", "q_ctg_name", "q_ctg_bgn", "q_ctg_end", "q_ctg_len", "aln_anchor_count", "src", "ctg", "ctg_bgn", "ctg_end", "orientation", "ctg_name"])]
    subs = []
    for sid, rgns in merged:
        ctg, src = seq_info[sid]
        src = src if src else "N/A"
        for b, e, _len, orientation, aln in rgns:
            aln = sorted(aln)
            q_bgn, q_end = aln[0][0][0], aln[-1][0][1]
            name = "%s::%s_%d_%d_%d" % (file_stem(src), ctg, b, e, orientation)
            if bed:
                lines.append("\t".join(map(str, [ctg, b, e, q_name, "#AAAAAA", orientation, q_len, len(aln), idx, src, q_bgn, q_end, name])))
            else:
                lines.append("\t".join(["%03d" % idx] + list(map(str, [q_name, q_bgn, q_end, q_len, len(aln), src, ctg, b, e, orientation, name]))))
            subs.append((name, sid, b, e, orientation))
    return "\n".join(lines) + "\n", subs
