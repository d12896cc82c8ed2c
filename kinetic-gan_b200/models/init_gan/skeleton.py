"""Skeleton-graph tables consumed by the networks (reference: models/init_gan/graph_ntu.py and
graph_h36m.py).  The reference derives them at ctor time with networkx (leaf-pruning coarsening,
graph_ntu.py:56-102).  The coarsening result is a constant of the skeleton, so here the per-level
bone lists / kept-joint maps / centres are frozen, and the derived quantities - hop distance,
column-normalised adjacency, the 3 "spatial" partitions (graph_ntu.py:117-171) and the up-sampling
neighbourhoods (graph_ntu.py:184-208) - are computed with vectorised numpy.  tests/test_graph.py checks
every table against the unmodified reference's output (tests/golden/graph_*.npz)."""
import numpy as np

FROZEN = {
    "ntu": dict(
        num_node=[25, 11, 5, 1], center=[20, 10, 4, 0],
        bones=[
            [[0, 1], [0, 12], [0, 16], [1, 20], [2, 20], [2, 3], [4, 20], [4, 5], [5, 6], [6, 7], [7, 21], [7, 22], [8, 20],
             [8, 9], [9, 10], [10, 11], [11, 23], [11, 24], [12, 13], [13, 14], [14, 15], [16, 17], [17, 18], [18, 19]],
            [[0, 10], [0, 6], [0, 8], [1, 10], [2, 10], [2, 3], [4, 10], [4, 5], [6, 7], [8, 9]],
            [[0, 4], [1, 4], [2, 4], [2, 3], [3, 4]],
            [],
        ],
        keep=[list(range(25)), [0, 2, 5, 7, 9, 11, 13, 14, 17, 18, 20], [2, 4, 6, 8, 10], [4]],
    ),
    "h36m": dict(
        num_node=[16, 7, 2, 1], center=[8, 3, 1, 0],
        bones=[
            [[0, 1], [0, 4], [0, 7], [1, 2], [2, 3], [4, 5], [5, 6], [7, 8], [8, 9], [8, 10], [8, 13], [10, 11], [11, 12],
             [13, 14], [14, 15]],
            [[0, 1], [0, 2], [0, 3], [3, 4], [3, 5], [3, 6]],
            [[0, 1]],
            [],
        ],
        keep=[list(range(16)), [0, 2, 5, 8, 9, 11, 14], [0, 3], [1]],
    ),
}


def _hop_distance(n, edge, max_hop):
    a = np.zeros((n, n))
    e = np.asarray(edge).reshape(-1, 2)
    a[e[:, 0], e[:, 1]] = 1
    a[e[:, 1], e[:, 0]] = 1
    hop = np.full((n, n), np.inf)
    reach = np.eye(n)
    layers = [reach > 0]
    for _ in range(max_hop):
        reach = reach @ a
        layers.append(reach > 0)
    for d in range(max_hop, -1, -1):
        hop[layers[d]] = d
    return hop


def _partitions(hop, center, max_hop, dilation):
    valid = list(range(0, max_hop + 1, dilation))
    adj = np.isin(hop, valid).astype(np.float64)
    deg = adj.sum(0)
    norm = adj / np.where(deg > 0, deg, 1.0)[None, :]
    dist = hop[:, center]                       # distance of every joint to the centre (inf beyond max_hop)
    dj, di = dist[:, None], dist[None, :]
    parts = []
    for h in valid:
        at = norm * (hop == h)
        root, close, further = at * (dj == di), at * (dj > di), at * (dj < di)
        if h == 0:
            parts.append(root)
        else:
            parts += [root + close, further]
    return np.stack(parts)


class SkeletonGraph:
    """Attribute surface of graph_ntu / Graph_h36m (graph_ntu.py:7-21): As, map, mapping, num_node, center,
    edge, nodes, hop_dis, lvls, max_hop, dilation."""

    def __init__(self, dataset, max_hop=1, dilation=1):
        f = FROZEN[dataset]
        self.max_hop, self.dilation, self.lvls = max_hop, dilation, 4
        self.num_node, self.center = list(f["num_node"]), list(f["center"])
        self.nodes = [np.arange(n) for n in self.num_node]
        self.map = [np.stack([np.arange(len(k)), np.asarray(k)], 1) for k in f["keep"]]
        self.edge = []
        for n, bones in zip(self.num_node, f["bones"]):
            self_link = [(i, i) for i in range(n)]
            self.edge.append(np.array(bones + [list(s) for s in self_link]) if bones else self_link)
        self.hop_dis = [_hop_distance(n, e, max_hop) for n, e in zip(self.num_node, self.edge)]
        self.As = [_partitions(h, c, max_hop, dilation) for h, c in zip(self.hop_dis, self.center)]
        self.mapping = self._upsample_hoods()

    def _upsample_hoods(self):
        """mapping[l]: for each level-l joint dropped at level l+1, [joint, coarse neighbours (level l+1 ids)...]."""
        out = []
        for l in range(self.lvls - 1):
            n = self.num_node[l]
            adj = np.zeros((n, n), bool)
            e = np.asarray(self.edge[l]).reshape(-1, 2)
            adj[e[:, 0], e[:, 1]] = adj[e[:, 1], e[:, 0]] = True
            kept = self.map[l + 1][:, 1]
            hoods = []
            for node in range(n):
                if node not in kept:
                    nb = np.nonzero(adj[node, kept])[0]
                    if len(nb):
                        hoods.append(np.concatenate([[node], nb]))
            out.append(hoods)
        return out

    def __str__(self):
        return str(self.As)
