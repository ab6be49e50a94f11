"""Cell-graph minimisation and isomorphism hash of an arch_vec (host logic of the sweep).

Same results as the reference's ``search_space.get_model_hash`` (search_space.py:21-29 -> graph_utils.py
get_model_graph_np :16-76 and graph_hash_np :145-180), re-implemented on plain Python lists:

* vertices: 0 = cell input, k = node k-1, n+1 = cell output; edge k-1 -> k carries node k-1's main op; the branch bits
  of node k (``arch_vec[k][1:]``, skip from vertex i into the node's output sum, model.py:16-22) become edges
  i -> k+2, i.e. into whoever consumes that sum;
* minimise: a ``zero`` vertex loses all its edges; vertices not on an input->output path are dropped;
* hash: the NAS-Bench-101 style iterated MD5 over (out-degree, in-degree, label) with labels -1 / op index / -2.
  Degrees are formatted as floats ("1.0") because the reference sums a float matrix -- the digest depends on it.

Known answers (reference README.md:61 and graph_utils.py:365-379): the default arch hashes to
36855332a5778e0df5114305bc3ce238; the 13 824 arch_vecs collapse to 8 242 distinct graphs.
"""
import hashlib

from . import search_space as ss


def get_model_graph(arch_vec, ops=None, minimize=True):
    """-> (adjacency matrix as list of float rows, vertex labels)."""
    ops = ss.all_ops if ops is None else ops
    n = len(arch_vec)
    size = n + 2
    adj = [[0.0] * size for _ in range(size)]
    labels = ['input'] + [ops[node[0]] for node in arch_vec] + ['output']
    for k in range(size - 1):
        adj[k][k + 1] = 1.0                       # main-op chain, and last node -> output
    for k, node in enumerate(arch_vec):
        for i, bit in enumerate(node[1:]):
            if bit:
                adj[i][k + 2] = 1.0               # skip i feeds the consumer of node k's sum
    if not minimize:
        return adj, labels
    for v, lab in enumerate(labels):
        if lab == 'zero':
            for u in range(size):
                adj[v][u] = adj[u][v] = 0.0

    def reach(src, forward):
        seen = {src}
        stack = [src]
        while stack:
            v = stack.pop()
            for u in range(size):
                if u not in seen and (adj[v][u] if forward else adj[u][v]):
                    seen.add(u)
                    stack.append(u)
        return seen

    alive = sorted(reach(0, True) & reach(size - 1, False))
    return [[adj[v][u] for u in alive] for v in alive], [labels[v] for v in alive]


def graph_hash(graph):
    adj, labels = graph
    n = len(adj)
    codes = ([-1] + [ss.all_ops.index(op) for op in labels[1:-1]] + [-2]) if labels else []
    assert len(codes) == n
    md5 = lambda text: hashlib.md5(text.encode('utf-8')).hexdigest()
    out_deg = [sum(row) for row in adj]
    in_deg = [sum(adj[v][u] for v in range(n)) for u in range(n)]
    hashes = [md5(str((out_deg[v], in_deg[v], codes[v]))) for v in range(n)]
    for _ in range(n):
        hashes = [md5(''.join(sorted(hashes[w] for w in range(n) if adj[w][v])) + '|' +
                      ''.join(sorted(hashes[w] for w in range(n) if adj[v][w])) + '|' + hashes[v]) for v in range(n)]
    return md5(str(sorted(hashes)))


def get_model_hash(arch_vec, ops=None, minimize=True):
    """search_space.py:21-29"""
    return graph_hash(get_model_graph(arch_vec, ops=ops, minimize=minimize))


def get_unique_architectures(ops=None, nodes=None):
    """First arch_vec of every isomorphism class, in enumeration order -> [(hash, arch_vec)] (8 242 for the default space)."""
    seen, out = set(), []
    for arch in ss.get_all_architectures(ops, nodes):
        h = get_model_hash(arch)
        if h not in seen:
            seen.add(h)
            out.append((h, arch))
    return out
