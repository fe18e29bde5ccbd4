"""bundles_oracle.py — TEST INFRASTRUCTURE ONLY (imported by tests/ only; the product path is pgr_tk_b200/csrc/bundles.cu).

Pure-Python restatement, statement by statement, of
  * graph_utils.rs:63-290  BiDiGraphWeightedDfs (new / move_to / next)
  * seq_db.rs:1013-1061    sort_adj_list_by_weighted_dfs
  * seq_db.rs:1063-1186    get_principal_bundles_from_adj_list
together with the third-party behaviour their results depend on, restated from the published sources because the
crates are not vendored under /root/reference:
  * petgraph 0.6.x `GraphMap<N, (), Directed>` (pgr-db/Cargo.toml: petgraph = "0.6.1"): `nodes: IndexMap<N, Vec<(N,
    CompactDirection)>>`; add_edge pushes (b, Outgoing) on a and (a, Incoming) on b (a != b), inserting unknown nodes in
    that order; neighbors_directed yields links whose direction matches OR whose endpoint is the node itself;
    remove_node = IndexMap::swap_remove of the node, then for each of its links Vec::swap_remove of the first matching
    back-link; `visit::Dfs` = LIFO stack, successors pushed in list order, visit on pop;
  * Rust std `BinaryHeap` (push = sift_up, pop = swap with last + sift_down_to_bottom + sift_up) with WeightedNode ordered
    by weight only; `slice::sort` is stable.
PARITY UNPINNED: the reference holds no test or fixture for these functions, and the Rust toolchain is absent, so this
restatement is pinned only to the reading above.  Vertices are (h0, h1, ori) tuples.
"""


def reverse(n):
    return (n[0], n[1], 1 - n[2])


OUT, IN = 0, 1


class GraphMap:
    """petgraph::graphmap::DiGraphMap<N, ()>"""

    def __init__(self):
        self.keys = []    # IndexMap order
        self.pos = {}     # node -> index in keys
        self.links = {}   # node -> [(neighbour, dir)]
        self.edges = set()

    def clone(self):
        g = GraphMap()
        g.keys = list(self.keys)
        g.pos = dict(self.pos)
        g.links = {k: list(v) for k, v in self.links.items()}
        g.edges = set(self.edges)
        return g

    def _entry(self, a):
        if a not in self.pos:
            self.pos[a] = len(self.keys)
            self.keys.append(a)
            self.links[a] = []
        return self.links[a]

    def add_edge(self, a, b):
        if (a, b) in self.edges:
            return
        self.edges.add((a, b))
        self._entry(a).append((b, OUT))
        if a != b:
            self._entry(b).append((a, IN))

    def nodes(self):
        return list(self.keys)

    def neighbors(self, a):
        return [n for n, d in self.links.get(a, []) if d == OUT]

    def neighbors_directed(self, a, direction):
        return [n for n, d in self.links.get(a, []) if d == direction or n == a]

    def all_edges(self):
        return list(self.edges)

    def remove_node(self, n):
        if n not in self.pos:
            return False
        i = self.pos.pop(n)
        last = self.keys.pop()
        if last != n:
            self.keys[i] = last
            self.pos[last] = i
        for succ, d in self.links.pop(n):
            edge = (n, succ) if d == OUT else (succ, n)
            sus = self.links.get(succ)
            if sus is not None:
                opp = IN if d == OUT else OUT
                for j, l in enumerate(sus):
                    if l == (n, opp):
                        sus[j] = sus[-1]
                        sus.pop()
                        break
            self.edges.discard(edge)
        return True


class BinaryHeap:
    """std::collections::BinaryHeap<WeightedNode<N>>; items are (weight, node), compared by weight only"""

    def __init__(self):
        self.d = []

    def _sift_up(self, start, pos):
        hole = self.d[pos]
        while pos > start:
            parent = (pos - 1) // 2
            if hole[0] <= self.d[parent][0]:
                break
            self.d[pos] = self.d[parent]
            pos = parent
        self.d[pos] = hole

    def _sift_down_to_bottom(self, pos):
        end = len(self.d)
        start = pos
        hole = self.d[pos]
        child = 2 * pos + 1
        while child <= max(end - 2, 0) and end >= 2:
            if self.d[child][0] <= self.d[child + 1][0]:
                child += 1
            self.d[pos] = self.d[child]
            pos = child
            child = 2 * pos + 1
        if child == end - 1:
            self.d[pos] = self.d[child]
            pos = child
        self.d[pos] = hole
        self._sift_up(start, pos)

    def push(self, item):
        self.d.append(item)
        self._sift_up(0, len(self.d) - 1)

    def pop(self):
        item = self.d.pop()
        if self.d:
            item, self.d[0] = self.d[0], item
            self._sift_down_to_bottom(0)
        return item

    def clear(self):
        self.d = []

    def is_empty(self):
        return not self.d


class BiDiGraphWeightedDfs:
    """graph_utils.rs:63-290"""

    def __init__(self, start, node_score):   # new(): graph_utils.rs:98-114
        self.priority_queue = BinaryHeap()
        self.discovered = set()
        self.next_node = None
        self.current_branch = 0
        self.branch_rank = 0
        self.global_rank = {}
        self.node_score = node_score
        s = node_score[start]                # .expect("Node not found")
        self.move_to(start)
        self.next_node = (s, start)
        self.global_rank[start] = 0

    def move_to(self, start):                # graph_utils.rs:158-165
        s = self.node_score[start]
        self.priority_queue.clear()
        self.priority_queue.push((s, start))
        self.next_node = (s, start)
        self.global_rank[start] = 0

    def next(self, graph):                   # graph_utils.rs:167-289
        branch = self.current_branch
        while True:
            if self.next_node is not None:
                node = self.next_node
                branch_rank = self.branch_rank
            else:
                if self.priority_queue.is_empty():
                    return None
                node = self.priority_queue.pop()
                self.branch_rank = 0
                branch_rank = 0
                self.current_branch += 1
                branch = self.current_branch
            v = node[1]
            if v in self.discovered:
                assert self.next_node is None, "the reference would loop forever here"
                continue
            self.discovered.add(v)
            rnode = reverse(v)
            self.discovered.add(rnode)
            f_out_count = 0
            succ_list_f = []
            for succ in graph.neighbors_directed(v, OUT):
                if v == succ or v == reverse(succ):
                    continue
                if succ not in self.discovered:
                    f_out_count += 1
                    succ_list_f.append((self.node_score[succ], succ))
            succ_list_r = []
            for succ in graph.neighbors_directed(rnode, OUT):
                if v == succ or v == reverse(succ):
                    continue
                if succ not in self.discovered:
                    succ_list_r.append((self.node_score[succ], succ))
            is_leaf = False
            if f_out_count == 0:
                is_leaf = True
                self.next_node = None
            if succ_list_f:
                succ_list_f.sort(key=lambda t: t[0])
                self.next_node = succ_list_f.pop()
                for s in succ_list_f:
                    self.priority_queue.push(s)
            if succ_list_r:
                succ_list_r.sort(key=lambda t: t[0])
                for s in succ_list_r:
                    self.priority_queue.push(s)
            node_rank = 0xFFFFFFFF
            p_node = None
            for n in graph.neighbors_directed(v, IN):
                r = self.global_rank.get(n)
                if r is not None and r < node_rank:
                    node_rank, p_node = r, n
            for n in graph.neighbors_directed(rnode, IN):
                r = self.global_rank.get(n)
                if r is not None and r < node_rank:
                    node_rank, p_node = r, n
            if node_rank == 0xFFFFFFFF:
                node_rank = 0
            node_rank += 1
            self.global_rank[v] = node_rank
            self.global_rank[rnode] = node_rank
            self.branch_rank += 1
            return (v, p_node, is_leaf, node_rank, branch, branch_rank)


def sort_adj_list_by_weighted_dfs(frag_count, adj_list, start):
    """seq_db.rs:1013-1061; frag_count[(h0, h1)] = frag_map.get(&key).unwrap().len(); adj_list = [(sid, v, w)]"""
    g = GraphMap()
    score = {}
    for _sid, v, w in adj_list:
        g.add_edge(v, w)
        score.setdefault(v, frag_count[(v[0], v[1])])
        score.setdefault(w, frag_count[(w[0], w[1])])
    walker = BiDiGraphWeightedDfs(start, score)
    out = []
    while True:
        r = walker.next(g)
        if r is None:
            break
        node, p_node, is_leaf, rank, branch_id, branch_rank = r
        out.append((node, p_node, score[node], is_leaf, rank, branch_id, branch_rank))
    return out


def get_principal_bundles_from_adj_list(frag_count, adj_list, path_len_cutoff):
    """seq_db.rs:1063-1186 -> (principal_bundles, filtered_adj_list)"""
    assert adj_list
    s = adj_list[0][1]
    sorted_adj_list = sort_adj_list_by_weighted_dfs(frag_count, adj_list, s)
    paths, path = [], []
    for v in sorted_adj_list:
        path.append(v[0])
        if v[3]:
            paths.append(path)
            path = []
    main = set()
    for p in paths:
        if len(p) > path_len_cutoff:
            for v in p:
                main.add((v[0], v[1]))
    g0 = GraphMap()
    filtered = []
    for sid, v, w in adj_list:
        if (v[0], v[1]) in main and (w[0], w[1]) in main:
            g0.add_edge(v, w)
            filtered.append((sid, v, w))
    g1 = g0.clone()
    terminal = set()
    for v, w in g0.all_edges():
        if len(g0.neighbors_directed(v, OUT)) > 1:
            terminal.add(v)
        if len(g0.neighbors_directed(w, IN)) > 1:
            terminal.add(v)
    starts = [v for v in g1.nodes() if len(g1.neighbors_directed(v, IN)) == 0]
    if not starts:
        ns = g1.nodes()
        if ns:
            starts.append(ns[0])
    bundles = []
    while starts:
        s = starts.pop()
        stack, discovered, path = [s], set(), []
        while True:                                   # petgraph::visit::Dfs::next
            v = None
            while stack:
                node = stack.pop()
                if node not in discovered:
                    discovered.add(node)
                    for succ in g1.neighbors(node):
                        if succ not in discovered:
                            stack.append(succ)
                    v = node
                    break
            if v is None:
                break
            path.append(v)
            if v in terminal:
                break
        if path:
            for v in path:
                g1.remove_node(v)
                g1.remove_node(reverse(v))
            starts = [v for v in g1.nodes() if len(g1.neighbors_directed(v, IN)) == 0]
            bundles.append(path)
        if not starts:
            ns = g1.nodes()
            if ns:
                starts.append(ns[0])
    bundles.sort(key=lambda b: -len(b))               # sort_by(b.len().cmp(a.len())), stable
    return bundles, filtered
