// bundles.cu — MAP-graph traversal above the adjacency list: sort_adj_list_by_weighted_dfs (seq_db.rs:1013-1061, driving
// graph_utils.rs:63-290 BiDiGraphWeightedDfs) and get_principal_bundles_from_adj_list (seq_db.rs:1063-1186).
//
// These are sequential graph walks on the host in the reference too; the GPU supplies the adjacency list
// (pgr_b200_adj_list) and the vertex weights (one batched key lookup in the device-resident index, weight_kernel).
//
// The reference's results depend on the iteration orders of petgraph 0.6 `DiGraphMap` (an insertion-ordered IndexMap of
// nodes, each with an insertion-ordered Vec of (neighbour, direction); `remove_node` = swap_remove of the node, and
// swap_remove of the back-links in the neighbours' lists), on `petgraph::visit::Dfs` (LIFO stack, successors pushed in
// list order) and on Rust's std BinaryHeap (sift_up / sift_down_to_bottom; WeightedNode compares by weight only).  OrdGraph
// and RustHeap below reproduce exactly those orders on integer vertex ids.  No reference test pins them (SURVEY §8c:
// "parity unpinned" for principal bundles); oracle/bundles_oracle.py is the independent restatement the tests compare with.
#include <algorithm>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "index.cuh"

namespace pgr {

struct GNode {
    uint64_t h0, h1;
    uint8_t ori;
    bool operator==(const GNode &o) const { return h0 == o.h0 && h1 == o.h1 && ori == o.ori; }
};
struct GNodeHash {
    size_t operator()(const GNode &n) const { return (size_t)(n.h0 * 0x9E3779B97F4A7C15ull ^ (n.h1 + n.ori) * 0xC2B2AE3D27D4EB4Full); }
};

// vertex ids in order of first appearance (a then b of every edge) = insertion order of DiGraphMap::add_edge
struct NodeTable {
    std::vector<GNode> nodes;
    std::unordered_map<GNode, uint32_t, GNodeHash> id;
    uint32_t intern(const GNode &n) {
        auto it = id.find(n);
        if (it != id.end()) return it->second;
        const uint32_t i = (uint32_t)nodes.size();
        nodes.push_back(n);
        id.emplace(n, i);
        return i;
    }
    int64_t find(const GNode &n) const {
        auto it = id.find(n);
        return it == id.end() ? -1 : (int64_t)it->second;
    }
};

enum : uint8_t { DIR_OUT = 0, DIR_IN = 1 };

// petgraph 0.6 GraphMap<N, (), Directed> restricted to what the two functions use
struct OrdGraph {
    struct Link { uint32_t n; uint8_t dir; };
    std::vector<std::vector<Link>> adj;   // by vertex id
    std::vector<int64_t> where;           // position in `order`, -1 = not (or no longer) a node
    std::vector<uint32_t> order;          // IndexMap order of the nodes
    std::unordered_set<uint64_t> edges;   // (a << 32 | b): add_edge of an existing edge only updates its weight

    explicit OrdGraph(size_t n_ids) : adj(n_ids), where(n_ids, -1) {}
    void touch(uint32_t a) {
        if (where[a] < 0) { where[a] = (int64_t)order.size(); order.push_back(a); }
    }
    void add_edge(uint32_t a, uint32_t b) {
        if (!edges.insert(((uint64_t)a << 32) | b).second) return;
        touch(a);
        adj[a].push_back({b, DIR_OUT});
        if (a != b) { touch(b); adj[b].push_back({a, DIR_IN}); }
    }
    bool has(uint32_t a) const { return where[a] >= 0; }
    // neighbors_directed(a, dir): links with that direction, plus self-loop links whatever their direction
    template <class F>
    void for_neighbors(uint32_t a, uint8_t dir, F f) const {
        if (!has(a)) return;
        for (const Link &l : adj[a]) if (l.dir == dir || l.n == a) f(l.n);
    }
    size_t count(uint32_t a, uint8_t dir) const {
        size_t c = 0;
        for_neighbors(a, dir, [&](uint32_t) { c++; });
        return c;
    }
    void remove_node(uint32_t n) {
        if (!has(n)) return;
        const size_t pos = (size_t)where[n];
        order[pos] = order.back();
        where[order[pos]] = (int64_t)pos;
        order.pop_back();
        where[n] = -1;
        std::vector<Link> links;
        links.swap(adj[n]);
        for (const Link &l : links) {
            const uint32_t a = (l.dir == DIR_OUT) ? n : l.n, b = (l.dir == DIR_OUT) ? l.n : n;
            edges.erase(((uint64_t)a << 32) | b);
            if (!has(l.n)) continue;   // remove_single_edge on a node that is gone (self loop)
            std::vector<Link> &sus = adj[l.n];
            const uint8_t opp = (l.dir == DIR_OUT) ? DIR_IN : DIR_OUT;
            for (size_t i = 0; i < sus.size(); i++)
                if (sus[i].n == n && sus[i].dir == opp) { sus[i] = sus.back(); sus.pop_back(); break; }
        }
    }
};

// Rust std::collections::BinaryHeap<WeightedNode>: max-heap on the weight, ties decided by the sift order
struct RustHeap {
    struct Item { uint32_t w, n; };
    std::vector<Item> d;
    void sift_up(size_t start, size_t pos) {
        const Item hole = d[pos];
        while (pos > start) {
            const size_t parent = (pos - 1) / 2;
            if (hole.w <= d[parent].w) break;
            d[pos] = d[parent];
            pos = parent;
        }
        d[pos] = hole;
    }
    void sift_down_to_bottom(size_t pos) {
        const size_t end = d.size(), start = pos;
        const Item hole = d[pos];
        size_t child = 2 * pos + 1;
        while (end >= 2 && child <= end - 2) {   // child <= end.saturating_sub(2)
            if (d[child].w <= d[child + 1].w) child++;
            d[pos] = d[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (end >= 1 && child == end - 1) { d[pos] = d[child]; pos = child; }
        d[pos] = hole;
        sift_up(start, pos);
    }
    void push(Item it) { d.push_back(it); sift_up(0, d.size() - 1); }
    Item pop() {
        Item it = d.back();
        d.pop_back();
        if (!d.empty()) { std::swap(it, d[0]); sift_down_to_bottom(0); }
        return it;
    }
    bool empty() const { return d.empty(); }
    void clear() { d.clear(); }
};

struct DfsRow { uint32_t node; int64_t prev; uint32_t weight; bool is_leaf; uint32_t rank, branch, branch_rank; };

// BiDiGraphWeightedDfs::new + next() until exhaustion (graph_utils.rs:98-114,167-289).  rev[v] = id of the reversed
// vertex or -1 when it is not a vertex of the graph.
static void weighted_dfs(const OrdGraph &g, const std::vector<int64_t> &rev, const std::vector<uint32_t> &score, uint32_t start,
                         std::vector<DfsRow> *out) {
    const size_t nv = g.adj.size();
    std::vector<uint8_t> discovered(nv, 0);
    std::vector<uint8_t> has_rank(nv, 0);
    std::vector<uint32_t> grank(nv, 0);
    RustHeap pq;
    // new(): empty(), move_to(start), next_node = start, global_rank[start] = 0
    pq.push({score[start], start});
    bool have_next = true;
    RustHeap::Item next_node = {score[start], start};
    has_rank[start] = 1; grank[start] = 0;
    uint32_t current_branch = 0, self_branch_rank = 0;
    for (;;) {
        // ---- one call of next() ----
        uint32_t branch_rank = 0, branch = current_branch;
        bool produced = false;
        for (;;) {
            RustHeap::Item node;
            if (have_next) { node = next_node; branch_rank = self_branch_rank; }
            else {
                if (pq.empty()) return;
                node = pq.pop();
                self_branch_rank = 0; branch_rank = 0;
                current_branch += 1; branch = current_branch;
            }
            const uint32_t v = node.n;
            if (discovered[v]) {
                // a discovered next_node would spin forever in the reference; it cannot happen (successors are checked)
                if (have_next) have_next = false;
                continue;
            }
            discovered[v] = 1;
            const int64_t rv = rev[v];
            if (rv >= 0) discovered[(size_t)rv] = 1;
            std::vector<RustHeap::Item> succ_f, succ_r;
            size_t f_out = 0;
            g.for_neighbors(v, DIR_OUT, [&](uint32_t s) {
                if (s == v || (int64_t)s == rv) return;   // no walk through self loops
                if (!discovered[s]) { f_out++; succ_f.push_back({score[s], s}); }
            });
            if (rv >= 0) g.for_neighbors((uint32_t)rv, DIR_OUT, [&](uint32_t s) {
                if (s == v || (int64_t)s == rv) return;
                if (!discovered[s]) succ_r.push_back({score[s], s});
            });
            bool is_leaf = false;
            if (f_out == 0) { is_leaf = true; have_next = false; }
            auto by_w = [](const RustHeap::Item &a, const RustHeap::Item &b) { return a.w < b.w; };
            if (!succ_f.empty()) {   // the same direction first
                std::stable_sort(succ_f.begin(), succ_f.end(), by_w);
                next_node = succ_f.back(); have_next = true;
                succ_f.pop_back();
                for (const auto &s : succ_f) pq.push(s);
            }
            if (!succ_r.empty()) {
                std::stable_sort(succ_r.begin(), succ_r.end(), by_w);
                for (const auto &s : succ_r) pq.push(s);
            }
            uint32_t node_rank = 0xFFFFFFFFu;
            int64_t p_node = -1;
            g.for_neighbors(v, DIR_IN, [&](uint32_t n) { if (has_rank[n] && grank[n] < node_rank) { node_rank = grank[n]; p_node = n; } });
            if (rv >= 0) g.for_neighbors((uint32_t)rv, DIR_IN, [&](uint32_t n) { if (has_rank[n] && grank[n] < node_rank) { node_rank = grank[n]; p_node = n; } });
            if (node_rank == 0xFFFFFFFFu) node_rank = 0;
            node_rank += 1;
            has_rank[v] = 1; grank[v] = node_rank;
            if (rv >= 0) { has_rank[(size_t)rv] = 1; grank[(size_t)rv] = node_rank; }
            self_branch_rank += 1;
            out->push_back({v, p_node, score[v], is_leaf, node_rank, branch, branch_rank});
            produced = true;
            break;
        }
        if (!produced) return;
    }
}

__global__ void weight_kernel(const SortKey *q, uint64_t n, const SortKey *ukeys, const uint64_t *offsets, uint64_t n_keys, uint32_t *w) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t kx = find_key(ukeys, n_keys, q[i].k0, q[i].k1);
    w[i] = kx < 0 ? 0xFFFFFFFFu : (uint32_t)(offsets[kx + 1] - offsets[kx]);   // frag_map.get(&key).unwrap().len()
}

// graph of an adjacency list + vertex weights from the index
struct AdjGraph {
    NodeTable tab;
    std::vector<std::pair<uint32_t, uint32_t>> edge_ids;   // per adjacency pair
    std::vector<int64_t> rev;
    std::vector<uint32_t> score;
};

static int build_adj_graph(pgr_b200_index *idx, const pgr_adj_pair *adj, size_t n_adj, AdjGraph *ag) {
    ag->edge_ids.resize(n_adj);
    for (size_t i = 0; i < n_adj; i++) {
        const uint32_t a = ag->tab.intern({adj[i].a0, adj[i].a1, adj[i].ori0});
        const uint32_t b = ag->tab.intern({adj[i].b0, adj[i].b1, adj[i].ori1});
        ag->edge_ids[i] = {a, b};
    }
    const size_t nv = ag->tab.nodes.size();
    ag->rev.assign(nv, -1);
    for (size_t v = 0; v < nv; v++) {
        GNode r = ag->tab.nodes[v];
        r.ori = (uint8_t)(1 - r.ori);
        ag->rev[v] = ag->tab.find(r);
    }
    // weights: one batched lookup of the vertex keys in the device-resident index
    PGR_CUDA(cudaSetDevice(idx->ctx->device));
    PGR_TRY(pgr_b200_index_finalize(idx));
    cudaStream_t st = idx->ctx->stream;
    std::vector<SortKey> keys(nv);
    for (size_t v = 0; v < nv; v++) keys[v] = {ag->tab.nodes[v].h0, ag->tab.nodes[v].h1};
    ag->score.assign(nv, 0);
    if (nv) {
        PGR_TRY(idx->scratch0.ensure(nv * sizeof(SortKey)));
        PGR_TRY(idx->scratch1.ensure(nv * sizeof(uint32_t)));
        PGR_CUDA(cudaMemcpyAsync(idx->scratch0.p, keys.data(), nv * sizeof(SortKey), cudaMemcpyHostToDevice, st));
        weight_kernel<<<(unsigned)ceil_div<uint64_t>(nv, 256), 256, 0, st>>>(idx->scratch0.as<SortKey>(), nv, idx->ukeys.as<SortKey>(),
                                                                              idx->offsets.as<uint64_t>(), idx->n_keys, idx->scratch1.as<uint32_t>());
        idx->launches += 1;
        PGR_CUDA(cudaGetLastError());
        PGR_CUDA(cudaMemcpyAsync(ag->score.data(), idx->scratch1.p, nv * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PGR_CUDA(cudaStreamSynchronize(st));
        for (size_t v = 0; v < nv; v++)
            if (ag->score[v] == 0xFFFFFFFFu) { set_error("adjacency vertex is not a key of the index (frag_map.get(..).unwrap(), seq_db.rs:1031)"); return PGR_E_ASSERT; }
    }
    return PGR_OK;
}

static pgr_graph_node to_c(const GNode &n) {
    pgr_graph_node c;
    memset(&c, 0, sizeof(c));
    c.h0 = n.h0; c.h1 = n.h1; c.ori = n.ori;
    return c;
}

}  // namespace pgr

using namespace pgr;

extern "C" {

int pgr_b200_sort_adj_list_by_weighted_dfs(pgr_b200_index *idx, const pgr_adj_pair *adj, size_t n_adj, const pgr_graph_node *start,
                                           pgr_dfs_node **out, size_t *n_out) {
    if (!idx || (!adj && n_adj) || !start || !out || !n_out) { set_error("NULL argument"); return PGR_E_ARG; }
    AdjGraph ag;
    PGR_TRY(build_adj_graph(idx, adj, n_adj, &ag));
    const int64_t s = ag.tab.find({start->h0, start->h1, start->ori});
    if (s < 0) { set_error("start vertex not in the adjacency list (\"Node not found\", graph_utils.rs:107)"); return PGR_E_ASSERT; }
    OrdGraph g(ag.tab.nodes.size());
    for (const auto &e : ag.edge_ids) g.add_edge(e.first, e.second);
    std::vector<DfsRow> rows;
    weighted_dfs(g, ag.rev, ag.score, (uint32_t)s, &rows);
    *out = (pgr_dfs_node *)result_alloc(std::max<size_t>(1, rows.size()) * sizeof(pgr_dfs_node));
    if (!*out) { set_error("out of host memory"); return PGR_E_ARG; }
    for (size_t i = 0; i < rows.size(); i++) {
        pgr_dfs_node &o = (*out)[i];
        memset(&o, 0, sizeof(o));
        o.node = to_c(ag.tab.nodes[rows[i].node]);
        o.has_prev = rows[i].prev >= 0;
        if (o.has_prev) o.prev = to_c(ag.tab.nodes[(size_t)rows[i].prev]);
        o.is_leaf = rows[i].is_leaf;
        o.weight = rows[i].weight; o.rank = rows[i].rank; o.branch = rows[i].branch; o.branch_rank = rows[i].branch_rank;
    }
    *n_out = rows.size();
    return PGR_OK;
}

int pgr_b200_principal_bundles(pgr_b200_index *idx, const pgr_adj_pair *adj, size_t n_adj, size_t path_len_cutoff, pgr_graph_node **vertices,
                               uint64_t **bundle_off, size_t *n_bundles, pgr_adj_pair **filtered, size_t *n_filtered) {
    if (!idx || (!adj && n_adj) || !vertices || !bundle_off || !n_bundles || !filtered || !n_filtered) { set_error("NULL argument"); return PGR_E_ARG; }
    if (n_adj == 0) { set_error("assert!(!adj_list.is_empty()) (seq_db.rs:1068)"); return PGR_E_ASSERT; }
    AdjGraph ag;
    PGR_TRY(build_adj_graph(idx, adj, n_adj, &ag));
    const size_t nv = ag.tab.nodes.size();
    // weighted DFS from adj_list[0].1, cut into paths at the leaves (seq_db.rs:1070-1086)
    std::vector<DfsRow> rows;
    {
        OrdGraph g(nv);
        for (const auto &e : ag.edge_ids) g.add_edge(e.first, e.second);
        weighted_dfs(g, ag.rev, ag.score, ag.edge_ids[0].first, &rows);
    }
    std::unordered_set<GNode, GNodeHash> main_vertices;   // (h0, h1) of the long paths; ori fixed to 0
    {
        size_t p0 = 0;
        for (size_t i = 0; i < rows.size(); i++) {
            if (!rows[i].is_leaf) continue;
            if (i + 1 - p0 > path_len_cutoff)
                for (size_t j = p0; j <= i; j++) { const GNode &v = ag.tab.nodes[rows[j].node]; main_vertices.insert({v.h0, v.h1, 0}); }
            p0 = i + 1;
        }
    }
    // g0 / filtered adjacency list (seq_db.rs:1100-1114)
    std::vector<pgr_adj_pair> flt;
    OrdGraph g1(nv);
    std::vector<std::pair<uint32_t, uint32_t>> g0_edges;   // all_edges(): distinct edges in insertion order
    for (size_t i = 0; i < n_adj; i++) {
        const GNode &v = ag.tab.nodes[ag.edge_ids[i].first], &w = ag.tab.nodes[ag.edge_ids[i].second];
        if (main_vertices.count({v.h0, v.h1, 0}) && main_vertices.count({w.h0, w.h1, 0})) {
            const size_t before = g1.edges.size();
            g1.add_edge(ag.edge_ids[i].first, ag.edge_ids[i].second);
            if (g1.edges.size() != before) g0_edges.push_back(ag.edge_ids[i]);
            flt.push_back(adj[i]);
        }
    }
    // terminal vertices on g0 (== g1 before any removal), seq_db.rs:1116-1126 (both conditions insert v)
    std::vector<uint8_t> terminal(nv, 0);
    for (const auto &e : g0_edges) {
        if (g1.count(e.first, DIR_OUT) > 1) terminal[e.first] = 1;
        if (g1.count(e.second, DIR_IN) > 1) terminal[e.first] = 1;
    }
    auto find_starts = [&](std::vector<uint32_t> &starts) {
        starts.clear();
        for (uint32_t v : g1.order) if (g1.count(v, DIR_IN) == 0) starts.push_back(v);
        if (starts.empty() && !g1.order.empty()) starts.push_back(g1.order[0]);   // the whole graph is a loop
    };
    std::vector<uint32_t> starts;
    find_starts(starts);
    std::vector<std::vector<uint32_t>> bundles;
    std::vector<uint8_t> seen(nv, 0);
    while (!starts.empty()) {
        const uint32_t s = starts.back();
        starts.pop_back();
        // petgraph::visit::Dfs from s until the first terminal vertex (inclusive)
        std::fill(seen.begin(), seen.end(), 0);
        std::vector<uint32_t> stack{s}, path;
        while (!stack.empty()) {
            const uint32_t v = stack.back();
            stack.pop_back();
            if (seen[v]) continue;
            seen[v] = 1;
            if (g1.has(v)) for (const auto &l : g1.adj[v]) if (l.dir == DIR_OUT && !seen[l.n]) stack.push_back(l.n);
            path.push_back(v);
            if (terminal[v]) break;
        }
        if (!path.empty()) {
            for (uint32_t v : path) {
                g1.remove_node(v);
                if (ag.rev[v] >= 0) g1.remove_node((uint32_t)ag.rev[v]);
            }
            find_starts(starts);   // starts.clear() + recount, then the loop fallback (seq_db.rs:1166-1180)
            bundles.push_back(path);
        } else if (starts.empty() && !g1.order.empty()) {
            starts.push_back(g1.order[0]);
        }
    }
    std::stable_sort(bundles.begin(), bundles.end(), [](const std::vector<uint32_t> &a, const std::vector<uint32_t> &b) { return a.size() > b.size(); });
    size_t total = 0;
    for (const auto &b : bundles) total += b.size();
    *vertices = (pgr_graph_node *)result_alloc(std::max<size_t>(1, total) * sizeof(pgr_graph_node));
    *bundle_off = (uint64_t *)result_alloc((bundles.size() + 1) * sizeof(uint64_t));
    *filtered = (pgr_adj_pair *)result_alloc(std::max<size_t>(1, flt.size()) * sizeof(pgr_adj_pair));
    if (!*vertices || !*bundle_off || !*filtered) { set_error("out of host memory"); return PGR_E_ARG; }
    size_t o = 0;
    for (size_t b = 0; b < bundles.size(); b++) {
        (*bundle_off)[b] = o;
        for (uint32_t v : bundles[b]) (*vertices)[o++] = to_c(ag.tab.nodes[v]);
    }
    (*bundle_off)[bundles.size()] = o;
    if (!flt.empty()) memcpy(*filtered, flt.data(), flt.size() * sizeof(pgr_adj_pair));
    *n_bundles = bundles.size();
    *n_filtered = flt.size();
    return PGR_OK;
}

}  // extern "C"
