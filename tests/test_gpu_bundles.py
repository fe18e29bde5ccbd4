"""GPU-side principal bundles (pgr_b200_sort_adj_list_by_weighted_dfs, pgr_b200_principal_bundles: adjacency list and
vertex weights from the device index, graph walks on the host as in the reference) against oracle/bundles_oracle.py.
Parity with the reference itself is unpinned for these functions (petgraph iteration orders)."""
import os
import sys

import numpy as np
import pytest

import pgr_tk_b200 as pg

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bundles_oracle as bo  # noqa: E402

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
COMP[[65, 67, 71, 84]] = [84, 71, 67, 65]


def repetitive_locus(seed, n_hap, n_units=6, unit_len=6000, flank=20000):
    """config-5-like haplotypes: units with copy number 1..4, 5 % inverted, SNPs at 2e-3"""
    rng = np.random.default_rng(seed)
    units = [ACGT[rng.integers(0, 4, size=unit_len)] for _ in range(n_units)]
    fl, fr = ACGT[rng.integers(0, 4, size=flank)], ACGT[rng.integers(0, 4, size=flank)]
    haps = []
    for h in range(n_hap):
        r = np.random.default_rng(seed * 100 + h)
        parts = [fl]
        for u in units:
            for _ in range(int(r.integers(1, 5))):
                v = u.copy()
                if r.random() < 0.05:
                    v = COMP[v][::-1].copy()
                parts.append(v)
        parts.append(fr)
        s = np.concatenate(parts)
        m = np.nonzero(r.random(len(s)) < 2e-3)[0]
        s[m] = ACGT[r.integers(0, 4, size=len(m))]
        haps.append(s.tobytes())
    return haps


def as_tuples(adj):
    return [(int(r["sid"]), (int(r["a0"]), int(r["a1"]), int(r["ori0"])), (int(r["b0"]), int(r["b1"]), int(r["ori1"]))) for r in adj]


def frag_counts(g):
    keys, offs, _ = g.export()
    k = keys.reshape(-1, 2)
    return {(int(k[i, 0]), int(k[i, 1])): int(offs[i + 1] - offs[i]) for i in range(len(k))}


def gnode(v):
    return (int(v["h0"]), int(v["h1"]), int(v["ori"]))


def check_index(g, min_counts=(0, 2), cutoffs=(0, 3, 8), keeps=None):
    cnt = frag_counts(g)
    n_checked = 0
    for mc in min_counts:
        adj = g.adj_list(mc, keeps)
        if adj.size == 0:
            continue
        adj_t = as_tuples(adj)
        start = adj_t[0][1]
        rows = g.sort_adj_list_by_weighted_dfs(adj, start)
        exp = bo.sort_adj_list_by_weighted_dfs(cnt, adj_t, start)
        assert len(rows) == len(exp)
        for r, e in zip(rows, exp):
            assert gnode(r["node"]) == e[0]
            assert (gnode(r["prev"]) if r["has_prev"] else None) == e[1]
            assert (int(r["weight"]), bool(r["is_leaf"]), int(r["rank"]), int(r["branch"]), int(r["branch_rank"])) == e[2:]
        for cutoff in cutoffs:
            bundles, flt = g.get_principal_bundles_from_adj_list(adj, cutoff)
            eb, ef = bo.get_principal_bundles_from_adj_list(cnt, adj_t, cutoff)
            assert [[gnode(v) for v in b] for b in bundles] == eb
            assert as_tuples(flt) == ef
            n_checked += len(eb)
    return n_checked


def test_repetitive_locus_bundles_match_oracle():
    haps = repetitive_locus(13, 12)
    g = pg.ShmmrIndex(pg.ShmmrSpec(48, 56, 4, 12), 0)
    g.add_batch(list(range(len(haps))), haps)
    assert check_index(g) > 3
    assert check_index(g, min_counts=(3,), cutoffs=(2,), keeps=[0, 5]) > 0
    pb = g.get_principal_bundles(0, 8)
    assert len(pb) > 0 and all(len(b) > 0 for b in pb)
    assert [len(b) for b in pb] == sorted((len(b) for b in pb), reverse=True)


def test_default_spec_pangenome_bundles_match_oracle():
    rng = np.random.default_rng(77)
    anc = ACGT[rng.integers(0, 4, size=150000)]
    haps = []
    for h in range(8):
        s = anc.copy()
        m = np.nonzero(rng.random(len(s)) < 2e-3)[0]
        s[m] = ACGT[rng.integers(0, 4, size=len(m))]
        if h % 3 == 0:
            a = int(rng.integers(30000, 60000))
            s = np.concatenate([s[:a], COMP[s[a:a + 9000]][::-1], s[a + 9000:]])
        haps.append(s.tobytes())
    g = pg.ShmmrIndex(pg.ShmmrSpec(80, 56, 4, 64), 0)
    g.add_batch(list(range(len(haps))), haps)
    assert check_index(g, min_counts=(0, 3), cutoffs=(0, 4)) > 0


def test_bundle_error_codes():
    haps = repetitive_locus(3, 3, n_units=2, unit_len=3000, flank=4000)
    g = pg.ShmmrIndex(pg.ShmmrSpec(48, 56, 4, 12), 0)
    g.add_batch([0, 1, 2], haps)
    adj = g.adj_list(0)
    with pytest.raises(pg.PgrError) as e:     # assert!(!adj_list.is_empty()), seq_db.rs:1068
        g.get_principal_bundles_from_adj_list(adj[:0], 0)
    assert e.value.code == -7
    with pytest.raises(pg.PgrError) as e:     # "Node not found", graph_utils.rs:107
        g.sort_adj_list_by_weighted_dfs(adj, (1, 2, 0))
    assert e.value.code == -7
    assert g.get_principal_bundles(10 ** 6, 0) == []   # ext.rs:499-501
