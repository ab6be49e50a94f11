"""Search-space helpers needed by the model builder and the sweep driver.

Mirrors the public names of nasbench_asr/search_space.py (all_ops :6, get_search_space :11-18,
get_all_architectures :32-47, arch_vec_to_names :77-93, get_model_hash :21-29 -> graph_utils.py).
"""
import itertools

all_ops = ['linear', 'conv5', 'conv5d2', 'conv7', 'conv7d2', 'zero']
ops_no_zero = all_ops[:-1]
default_nodes = 3


def get_search_space(ops=None, nodes=None):
    ops = all_ops if ops is None else ops
    nodes = default_nodes if nodes is None else nodes
    return [[len(ops)] + [2] * (n + 1) for n in range(nodes)]


def get_all_architectures(ops=None, nodes=None):
    """Yield every arch_vec; the FIRST entry varies fastest, like the reference's odometer."""
    space = get_search_space(ops, nodes)
    flat = [d for node in space for d in node]
    for combo in itertools.product(*[range(d) for d in reversed(flat)]):
        vals = list(reversed(combo))
        out, i = [], 0
        for node in space:
            out.append(vals[i:i + len(node)])
            i += len(node)
        yield out


def arch_vec_to_names(arch_vec, ops=None):
    # NB: the reference ignores `ops` and always indexes the global table (search_space.py:93).
    return [[all_ops[node[0]]] + list(node[1:]) for node in arch_vec]


def get_model_hash(arch_vec, ops=None, minimize=True):
    """search_space.py:21-29"""
    from .graph_utils import get_model_hash as _h
    return _h(arch_vec, ops=ops, minimize=minimize)


def validate_arch(arch_vec):
    if len(arch_vec) != default_nodes:
        raise ValueError(f'expected {default_nodes} nodes, got {len(arch_vec)}')
    for n, node in enumerate(arch_vec):
        if len(node) != n + 2:
            raise ValueError(f'node {n} needs {n + 2} entries, got {node}')
        if not 0 <= node[0] < len(all_ops):
            raise ValueError(f'Operation id {node[0]} is not implemented')
        if any(b not in (0, 1) for b in node[1:]):
            raise ValueError(f'Invalid branch operations: {node[1:]}, expected a vector of 0 (no skip-con.) '
                             'and 1 (skip-con. present)')
