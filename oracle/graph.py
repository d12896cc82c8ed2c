"""Oracle restatement of the skeleton-graph tables (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference/models/init_gan/graph_ntu.py and graph_h36m.py:
  get_edge            graph_ntu.py:25-114   (level-0 bones, 3x leaf-pruning coarsening)
  get_hop_distance    graph_ntu.py:148-160
  normalize_digraph   graph_ntu.py:163-171  (column normalisation A . D^-1)
  get_adjacency       graph_ntu.py:117-144  (ST-GCN "spatial" 3-partition)
  upsample_mapping    graph_ntu.py:184-208  (neighbourhoods of re-inserted joints, reversed at :21)

The reference drives the coarsening with networkx; networkx is a third-party dependency that is
not part of /root/reference (requirements.txt:17 pins networkx==2.5).  Its behaviour that matters
here - insertion-ordered node/adjacency iteration and `cycle_basis` - is restated below in plain
Python.  Pinned against tables dumped from the unmodified reference: tests/golden/graph_*.npz.
"""
import numpy as np

NTU_BONES_1BASED = [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9),
                    (11, 10), (12, 11), (1, 13), (14, 13), (15, 14), (16, 15), (1, 17), (18, 17), (19, 18),
                    (20, 19), (22, 8), (23, 8), (24, 12), (25, 12)]          # graph_ntu.py:33-37
H36M_BONES = [(1, 2), (2, 3), (0, 1), (4, 5), (5, 6), (0, 4), (0, 7), (7, 8), (8, 9), (8, 10), (10, 11),
              (11, 12), (8, 13), (13, 14), (14, 15)]                         # graph_h36m.py:33-37


class _OrderedGraph:
    """Undirected simple graph with insertion-ordered nodes and adjacency (what networkx.Graph gives)."""

    def __init__(self, n):
        self.adj = {i: {} for i in range(n)}

    def add_edge(self, u, v):
        self.adj[u][v] = True
        self.adj[v][u] = True

    def remove_node(self, u):
        for v in list(self.adj[u]):
            del self.adj[v][u]
        del self.adj[u]

    def nodes(self):
        return list(self.adj)

    def edges_of(self, u):
        return [(u, v) for v in self.adj[u]]

    def edges(self):
        seen, out = set(), []
        for u in self.adj:
            for v in self.adj[u]:
                if v not in seen:
                    out.append((u, v))
            seen.add(u)
        return out

    def relabel_in_order(self):
        m = {u: i for i, u in enumerate(self.adj)}
        g = _OrderedGraph(0)
        g.adj = {m[u]: {m[v]: True for v in nb} for u, nb in self.adj.items()}
        return g


def _first_basis_cycle_len(g):
    """Length of the first cycle networkx.cycle_basis would report (0 if the graph is a forest)."""
    gnodes = dict.fromkeys(g.adj)
    while gnodes:
        root = gnodes.popitem()[0]
        stack, pred, used = [root], {root: root}, {root: set()}
        while stack:
            z = stack.pop()
            zused = used[z]
            for nbr in g.adj[z]:
                if nbr not in used:
                    pred[nbr] = z
                    stack.append(nbr)
                    used[nbr] = {z}
                elif nbr == z:
                    return 1
                elif nbr not in zused:
                    pn = used[nbr]
                    cyc = [nbr, z]
                    p = pred[z]
                    while p not in pn:
                        cyc.append(p)
                        p = pred[p]
                    cyc.append(p)
                    return len(cyc)
        for node in pred:
            gnodes.pop(node, None)
    return 0


def coarsen(num_node, bones, center0, lvls=4, skip=None):
    """graph_ntu.py:25-114.  Returns per-level (num_node, edge list incl. self links, map, center)."""
    g = _OrderedGraph(num_node)
    for u, v in bones:
        g.add_edge(u, v)
    maps = [np.array([[i, x] for i, x in enumerate(g.nodes())])]
    edges = [np.array(g.edges() + [(i, i) for i in g.nodes()])]
    nums, centers = [num_node], [center0]
    for it in range(lvls - 1):
        stay, start = [], 1
        while True:
            remove = []
            for i in g.nodes():
                if len(g.adj[i]) == start and i not in stay:
                    if skip is not None and skip(i, it):
                        continue
                    lost = [k for _, k in g.edges_of(i)]
                    stay.extend(lost)
                    for a in lost:
                        for b in lost:
                            if a != b:
                                g.add_edge(a, b)
                    remove.append(i)
            if start > 10:
                break
            for i in remove:
                g.remove_node(i)
            if len(g.adj) and _first_basis_cycle_len(g) == len(g.adj):
                for x in [x for x in g.nodes() if x not in stay]:
                    g.remove_node(x)
            start += 1
        maps.append(np.array([[i, x] for i, x in enumerate(g.nodes())]))
        for i, x in enumerate(g.nodes()):
            if x == centers[-1]:
                centers.append(i)
        g = g.relabel_in_order()
        e = g.edges()
        self_link = [(i, i) for i in g.nodes()]
        edges.append(np.array(e + self_link) if len(e) else np.array(self_link))
        nums.append(len(g.adj))
    return nums, edges, maps, centers


def hop_distance(n, edge, max_hop=1):
    """graph_ntu.py:148-160."""
    a = np.zeros((n, n))
    for i, j in edge:
        a[j, i] = 1
        a[i, j] = 1
    hop = np.full((n, n), np.inf)
    arrive = [np.linalg.matrix_power(a, d) > 0 for d in range(max_hop + 1)]
    for d in range(max_hop, -1, -1):
        hop[arrive[d]] = d
    return hop


def spatial_partitions(n, edge, center, max_hop=1, dilation=1):
    """graph_ntu.py:117-144 + :163-171.  Returns (K=3, n, n) float64."""
    hop = hop_distance(n, edge, max_hop)
    valid = range(0, max_hop + 1, dilation)
    adj = np.zeros((n, n))
    for h in valid:
        adj[hop == h] = 1
    deg = adj.sum(0)
    norm = adj * np.where(deg > 0, 1.0 / np.where(deg > 0, deg, 1), 0.0)[None, :]
    parts = []
    for h in valid:
        root, close, further = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        for i in range(n):
            for j in range(n):
                if hop[j, i] == h:
                    if hop[j, center] == hop[i, center]:
                        root[j, i] = norm[j, i]
                    elif hop[j, center] > hop[i, center]:
                        close[j, i] = norm[j, i]
                    else:
                        further[j, i] = norm[j, i]
        if h == 0:
            parts.append(root)
        else:
            parts.append(root + close)
            parts.append(further)
    return np.stack(parts)


def upsample_neighbourhoods(maps, nums, edges, lvls=4):
    """graph_ntu.py:184-208, already reversed as at graph_ntu.py:21: result[l] lifts level l+1 -> l."""
    hoods = []
    for i in range(lvls - 1, 0, -1):
        n = i - 1
        elist = {(int(a), int(b)) for a, b in np.asarray(edges[n]).tolist()}
        kept = maps[i][:, 1].tolist()
        level = []
        for node in range(nums[n]):
            if node not in kept:
                hood = [int(c[0]) for c in maps[i] if (node, int(c[1])) in elist or (int(c[1]), node) in elist]
                if hood:
                    level.append(np.array([node] + hood))
        hoods.append(level)
    return hoods[::-1]


class SkeletonTables:
    """Everything the networks consume from graph_ntu / Graph_h36m (graph_ntu.py:7-21)."""

    def __init__(self, dataset="ntu"):
        if dataset == "ntu":
            bones = [(i - 1, j - 1) for i, j in NTU_BONES_1BASED]
            nums, edges, maps, centers = coarsen(25, bones, 20)
        else:
            # graph_h36m.py:60 keeps joint 9 (head) during the first coarsening
            nums, edges, maps, centers = coarsen(16, H36M_BONES, 8, skip=lambda i, it: i == 9 and it == 0)
        self.lvls = 4
        self.num_node, self.edge, self.map, self.center = nums, edges, maps, centers
        self.As = [spatial_partitions(nums[l], edges[l], centers[l]) for l in range(self.lvls)]
        self.mapping = upsample_neighbourhoods(maps, nums, edges, self.lvls)
