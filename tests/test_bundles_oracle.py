"""CPU checks of oracle/bundles_oracle.py (the restatement the GPU-side bundle tests compare with): hand-derived cases
of the weighted DFS (graph_utils.rs:167-289) and of the principal-bundle decomposition (seq_db.rs:1063-1186), the Rust
BinaryHeap order, and structural invariants on random bidirected graphs.  Parity with the reference itself is unpinned
(no fixture exists)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bundles_oracle as bo  # noqa: E402


def bidirected(edges):
    """the adjacency list as frag_map_to_adj_list emits it: (sid, v, w) and (sid, rev(w), rev(v)) (seq_db.rs:921-943)"""
    out = []
    for v, w in edges:
        out.append((0, v, w))
        out.append((0, bo.reverse(w), bo.reverse(v)))
    return out


def node(i, ori=0):
    return (i, i + 1000, ori)


def test_binary_heap_matches_rust_order():
    # max-heap by weight; equal weights come out in the order Rust's sift_down_to_bottom/sift_up leaves them
    h = bo.BinaryHeap()
    for w, n in [(3, "a"), (5, "b"), (5, "c"), (1, "d"), (5, "e"), (3, "f")]:
        h.push((w, n))
    out = [h.pop() for _ in range(6)]
    assert [w for w, _ in out] == [5, 5, 5, 3, 3, 1]
    # hand-run of std's algorithm: after the pushes the vector is [b5, e5, c5, d1, a3, f3]; pop 1 moves f3 to the root and
    # sift_down_to_bottom takes the RIGHT child on the e5 <= c5 tie -> [c5, e5, f3, d1, a3]; and so on
    assert [n for _, n in out] == ["b", "c", "e", "f", "a", "d"]


def test_linear_chain_is_one_bundle():
    chain = [node(i) for i in range(6)]
    adj = bidirected(list(zip(chain[:-1], chain[1:])))
    cnt = {(n[0], n[1]): 4 for n in chain}
    rows = bo.sort_adj_list_by_weighted_dfs(cnt, adj, chain[0])
    assert [r[0] for r in rows] == chain                       # walks forward along the chain
    assert [r[3] for r in rows] == [False] * 5 + [True]        # only the end is a leaf
    assert [r[4] for r in rows] == [1, 2, 3, 4, 5, 6]          # global rank = 1 + min rank of the predecessors
    assert [r[1] for r in rows] == [None] + chain[:-1]
    assert all(r[5] == 0 for r in rows) and [r[6] for r in rows] == list(range(6))
    bundles, flt = bo.get_principal_bundles_from_adj_list(cnt, adj, 0)
    assert len(flt) == len(adj)
    assert len(bundles) == 1 and len(bundles[0]) == 6
    assert {(v[0], v[1]) for v in bundles[0]} == {(n[0], n[1]) for n in chain}
    # with a cutoff longer than the only path nothing is left
    bundles, flt = bo.get_principal_bundles_from_adj_list(cnt, adj, 6)
    assert bundles == [] and flt == []


def test_branch_prefers_heavier_successor_and_splits_bundles():
    a, b, c, d, e, f = (node(i) for i in range(6))
    # a -> b -> {c (weight 9), d (weight 2)}; c -> e; d -> f
    adj = bidirected([(a, b), (b, c), (b, d), (c, e), (d, f)])
    cnt = {(n[0], n[1]): 5 for n in (a, b, e, f)}
    cnt[(c[0], c[1])] = 9
    cnt[(d[0], d[1])] = 2
    rows = bo.sort_adj_list_by_weighted_dfs(cnt, adj, a)
    order = [r[0] for r in rows]
    assert order[:4] == [a, b, c, e]           # heavier branch first
    assert set(order[4:]) == {d, f} and order.index(d) < order.index(f)
    assert rows[3][3] and rows[-1][3]          # e and f are leaves
    bundles, _ = bo.get_principal_bundles_from_adj_list(cnt, adj, 0)
    keys = [sorted((v[0], v[1]) for v in bd) for bd in bundles]
    assert sorted(map(len, bundles), reverse=True) == [len(bd) for bd in bundles]
    # every vertex key appears in exactly one bundle
    flat = [k for ks in keys for k in ks]
    assert len(flat) == len(set(flat)) == 6


def test_random_graph_invariants():
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(5, 40))
        nodes = [node(i, int(rng.integers(0, 2))) for i in range(n)]
        edges = set()
        for i in range(n - 1):
            if rng.random() < 0.85:
                edges.add((nodes[i], nodes[i + 1]))
        for _ in range(int(rng.integers(0, n))):
            i, j = (int(x) for x in rng.integers(0, n, size=2))
            if i != j:
                edges.add((nodes[i], nodes[j]))
        if not edges:
            continue
        adj = bidirected(sorted(edges))
        cnt = {(nd[0], nd[1]): int(rng.integers(1, 6)) for nd in nodes}
        rows = bo.sort_adj_list_by_weighted_dfs(cnt, adj, adj[0][1])
        seen = [(r[0][0], r[0][1]) for r in rows]
        assert len(seen) == len(set(seen))                        # a vertex and its reverse are visited once
        succ = {}
        for _, v, w in adj:
            succ.setdefault(v, set()).add(w)
        for cutoff in (0, 2):
            bundles, flt = bo.get_principal_bundles_from_adj_list(cnt, adj, cutoff)
            flat = [(v[0], v[1]) for bd in bundles for v in bd]
            assert len(flat) == len(set(flat))
            fl = {(v, w) for _, v, w in flt}
            assert fl <= {(v, w) for _, v, w in adj}
            assert [len(b) for b in bundles] == sorted((len(b) for b in bundles), reverse=True)
