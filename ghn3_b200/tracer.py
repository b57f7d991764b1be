"""
Host-side tracer: nn.Module -> computational graph (op ids, 1-hop edges, node_info), behaviour-compatible with the
reference's Graph._build_graph / _filter_graph / _construct_features (ghn3/graph.py:392-753, 800-908). The graph
TRACING stays on the host (BASELINE.json north_star); only the shortest-path "virtual edges" moved to the GPU.

The algorithm follows the reference step by step, because node ORDER and names are part of the contract (the node ->
parameter mapping of nn.py:594-692 and every downstream index depend on them):
  1. run the model once, walk the autograd graph from the output (pre-order DFS over grad_fn.next_functions); an op
     whose inputs are parameters is represented by its parameter leaves (graph.py:423-480)
  2. drop unsupported / redundant ops, re-wiring their predecessors to their successors (graph.py:648-753)
  3. repair weight / softmax edge directions, Swin special cases (graph.py:511-598)
  4. add the input node, topological sort (networkx), ViT pos-enc and SqueezeNet fix-ups (graph.py:604-641)
  5. node features and node_info (graph.py:800-908)
It is written around adjacency *sets* instead of dense-matrix scans, so a trace costs the model's forward pass plus
O(N * degree) host work. Verified bit-exact (ops, edges, node_info) against graphs produced by the reference for the
torchvision classification models (tests/golden/graphs_tv.json.gz).
"""
import copy
import os

import networkx as nx
import numpy as np
import torch
import torch.nn as nn
import torchvision.models as tvm

from .graph import PRIMITIVES_DEEPNETS1M

_PRIM_ID = {op: i for i, op in enumerate(PRIMITIVES_DEEPNETS1M)}


def _conv_name(module, op_name):
    if 'bias' in op_name:
        return 'bias'
    if isinstance(module, nn.Conv2d) and module.groups > 1:
        return 'dil_conv' if min(module.dilation) > 1 else 'sep_conv'
    return 'conv'


def _module_table():
    """Supported module types -> primitive name (reference MODULES, graph.py:1114-1138)."""
    table = {
        nn.Conv2d: _conv_name,
        nn.Linear: _conv_name,
        nn.modules.linear.NonDynamicallyQuantizableLinear: _conv_name,
        nn.modules.activation.MultiheadAttention: _conv_name,
        nn.BatchNorm2d: lambda m, n: 'bn',
        nn.LayerNorm: lambda m, n: 'ln',
        tvm.convnext.LayerNorm2d: lambda m, n: 'ln',
        nn.modules.sparse.Embedding: lambda m, n: 'pos_enc',
        tvm.vision_transformer.Encoder: lambda m, n: 'pos_enc',
    }
    try:                                   # Hugging Face GPT-2 style layers, only if transformers is already loaded
        import sys
        if 'transformers' in sys.modules:
            from transformers.pytorch_utils import Conv1D
            table[Conv1D] = _conv_name
    except Exception:
        pass
    return table


_OP_TABLE = {'input': 'input', 'Mean': 'glob_avg', 'AdaptiveAvgPool2D': 'glob_avg', 'MaxPool2DWithIndices': 'max_pool',
             'AvgPool2D': 'avg_pool', 'Softmax': 'msa', 'Mul': 'cse', 'Add': 'sum', 'Cat': 'concat',
             'skip_connect': 'sum'}


class _Node:
    __slots__ = ('name', 'module', 'size', 'ksize', 'is_op')

    def __init__(self, name, module=None, size=None, ksize=None, is_op=True):
        self.name, self.module, self.size, self.ksize, self.is_op = name, module, size, ksize, is_op


def _cell_index(name, n_cells):
    pos = name.find('cells.')
    if pos >= 0:
        return int(name[pos + 6:].split('.', 1)[0])
    if name.startswith(('classifier', 'auxiliary')):
        return n_cells - 1
    if n_cells == 1 or name.startswith(('stem', 'pos_enc')):
        return 0
    return None


# ----------------------------------------------------------------------------------------------------------------
# 1. autograd walk
# ----------------------------------------------------------------------------------------------------------------
def _forward_for_graph(model, input_sz):
    """Runs the model once to obtain the autograd graph. First choice: a META forward -- parameters and buffers are
    replaced by meta tensors through torch.func.functional_call, so no arithmetic is done and the model (weights,
    BatchNorm statistics) is left untouched; the graph and every size are the same. Models whose forward needs real
    values (data-dependent control flow) fall back to the real forward of the reference (graph.py:400-420).
    Returns (outputs, owner: id(leaf tensor) -> (parameter name, module))."""
    mods = dict(model.named_modules())
    if not hasattr(model, 'get_var') and os.environ.get('GHN3_TRACE_META', '1') != '0':
        try:
            meta, owner, seen = {}, {}, {}
            for mod_name, mod in mods.items():
                for p_name, p in mod.named_parameters(recurse=False):
                    if p is None:
                        continue
                    full = (mod_name + '.' if mod_name else '') + p_name
                    if id(p) not in seen:                                  # shared parameters stay shared
                        seen[id(p)] = torch.empty_like(p, device='meta', requires_grad=True)
                        owner[id(seen[id(p)])] = (mod_name + '.' + p_name, mod)
                    meta[full] = seen[id(p)]
            for b_name, b in model.named_buffers():
                meta[b_name] = torch.empty_like(b, device='meta')
            with torch.enable_grad():
                out = torch.func.functional_call(model, meta, (torch.empty(2, *input_sz, device='meta'),))
            return out, owner, list(seen.values())
        except Exception:                                                  # noqa: BLE001 -- any meta-kernel gap
            pass
    owner = {}
    for mod_name, mod in mods.items():
        for p_name, p in mod.named_parameters(recurse=False):
            if p is not None and id(p) not in owner:
                owner[id(p)] = (mod_name + '.' + p_name, mod)
    device = next(model.parameters()).device
    with torch.enable_grad():
        out = model.get_var() if hasattr(model, 'get_var') else model(torch.randn(2, *input_sz, device=device))
    return out, owner, None


def _autograd_graph(model, input_sz):
    out, owner, _keepalive = _forward_for_graph(model, input_sz)
    if isinstance(out, dict):
        out = list(out.values())
    if not isinstance(out, (tuple, list)):
        out = [out]

    order, nodes, seen, edges = [], {}, {}, []     # order: node keys in creation order

    def open_fn(fn):
        """Creates the node(s) of fn (pre-order) and returns (first key, last key, type name)."""
        tname = type(fn).__name__
        first = last = None
        if 'AccumulateGrad' not in tname:
            leaves = []
            for nxt in fn.next_functions:
                cand = nxt[0]
                if cand is not None and hasattr(cand, 'variable'):
                    leaves.append(cand)
            if leaves:
                for leaf in leaves:
                    var = leaf.variable
                    pname, mod = owner[id(var)]
                    key = id(leaf)
                    if first is None:
                        first = key
                    last = key
                    seen[leaf] = (key, pname)
                    if key not in nodes:
                        order.append(key)
                    nodes[key] = _Node(pname, mod, tuple(var.size()), None, is_op=False)
            else:
                key = id(fn)
                first = last = key
                ks = getattr(fn, '_saved_kernel_size', None)
                if key not in nodes:
                    order.append(key)
                nodes[key] = _Node(tname, None, None, None if ks is None else str(ks))
        seen[fn] = (last, tname)
        return first, last, tname

    for v in out:
        if v is None or v.grad_fn is None:
            continue
        root = v.grad_fn
        if root in seen:
            continue
        first, last, tname = open_fn(root)
        stack = [(root, first, iter(root.next_functions))]
        while stack:
            fn, link_start, it = stack[-1]
            advanced = False
            for nxt in it:
                child = nxt[0]
                if child is None:
                    continue
                if child in seen:
                    c_link, c_name = seen[child]
                else:
                    c_first, c_last, c_name = open_fn(child)
                    stack.append((child, c_first, iter(child.next_functions)))
                    advanced = True
                    break                                   # recurse; the edge is added when the child returns
                if c_link is not None and link_start != c_link:
                    edges.append((link_start, c_link) if 'bias' in c_name else (c_link, link_start))
            if advanced:
                continue
            stack.pop()
            if stack:                                       # "return" of the child to its parent
                p_fn, p_start, _ = stack[-1]
                c_link, c_name = seen[fn]
                if c_link is not None and p_start != c_link:
                    edges.append((p_start, c_link) if 'bias' in c_name else (c_link, p_start))
    index = {k: i for i, k in enumerate(order)}
    node_list = [nodes[k] for k in order]
    n = len(node_list)
    succ = [set() for _ in range(n)]
    pred = [set() for _ in range(n)]
    for a, b in edges:
        ia, ib = index[a], index[b]
        succ[ia].add(ib)
        pred[ib].add(ia)
    return node_list, succ, pred


# ----------------------------------------------------------------------------------------------------------------
# 2. filtering
# ----------------------------------------------------------------------------------------------------------------
def _filter(nodes, succ, pred, table, only=None):
    if only is None:
        unsupported = set()
        for nd in nodes:
            cut = nd.name.find('Backward')
            op = nd.name if cut == -1 else nd.name[:cut]
            ok = False
            if 'norm' in type(nd.module).__name__.lower() and op.endswith('.bias'):
                pass                  # biases of norm layers are predicted but are not graph nodes (graph.py:666-671)
            else:
                ok = any(isinstance(nd.module, t) for t in table)
            if not ok and op not in _OP_TABLE:
                unsupported.add(nd.name)
        passes = ['Mul'] + sorted(unsupported) + ['Mean', 'Add', 'Cat']
    else:
        passes = list(only)
    has_cse = any(('sigmoid' in nd.name.lower()) or ('swish' in nd.name.lower()) for nd in nodes)
    n_in = [len(p) for p in pred]
    for pat in passes:
        n = len(nodes)
        keep_mask = [True] * n
        for i, nd in enumerate(nodes):
            if pat not in nd.name:
                continue
            neigh = None
            try:
                neigh = {j: nodes[i + j].name.lower() for j in (-1, -2, -3, 1)}
                head = any(neigh[j].startswith(('classifier', 'fc', 'head')) for j in (-1, -2))
            except IndexError:
                head = True
            keep = True
            if nd.name.startswith('Mean'):
                if has_cse:
                    keep = head
            elif nd.name.startswith('Mul'):
                keep = has_cse and not head and (neigh[-2].startswith(('hard', 'sigmoid')) or
                                                 neigh[-3].startswith(('relu', 'mean')) or
                                                 neigh[1].startswith(('hard', 'sigmoid', 'relu')))
            elif nd.name.startswith(('Cat', 'Add')):
                keep = n_in[i] > 1
            else:
                keep = False
            if not keep:
                keep_mask[i] = False
                for n1 in list(succ[i]):
                    for n2 in list(pred[i]):
                        if n1 != n2:
                            succ[n2].add(n1)
                            pred[n1].add(n2)
        if not all(keep_mask):
            remap = {}
            for i, k in enumerate(keep_mask):
                if k:
                    remap[i] = len(remap)
            nodes = [nd for nd, k in zip(nodes, keep_mask) if k]
            n_in = [v for v, k in zip(n_in, keep_mask) if k]
            succ = [{remap[j] for j in succ[i] if j in remap} for i in range(n) if keep_mask[i]]
            pred = [{remap[j] for j in pred[i] if j in remap} for i in range(n) if keep_mask[i]]
    return nodes, succ, pred


def _set_edge(succ, pred, a, b, on):
    if on:
        succ[a].add(b)
        pred[b].add(a)
    else:
        succ[a].discard(b)
        pred[b].discard(a)


# ----------------------------------------------------------------------------------------------------------------
# 3. edge repairs
# ----------------------------------------------------------------------------------------------------------------
def _fix_weight_edges(nodes, succ, pred):
    """graph.py:511-551: a weight that hangs off its layer as a source node is moved in front of the layer's bias."""
    for i in range(len(nodes)):
        node = nodes[i]
        if pred[i] or 'weight' not in node.name:
            continue
        for out in sorted(succ[i]):
            same_layer = node.module == nodes[out].module
            qkv = len(pred[i]) == 0 and 'softmax' in nodes[out].name.lower()
            if not (same_layer or qkv):
                continue
            n_out = len(succ[i])
            others = sorted(pred[out] - {i})
            if not others:
                continue
            nodes[i], nodes[out] = nodes[out], nodes[i]
            _set_edge(succ, pred, i, out, False)
            _set_edge(succ, pred, out, i, True)
            if n_out == 1:
                moved = sorted(succ[out] - {i})
                if not moved:
                    continue
                for t in moved:
                    _set_edge(succ, pred, out, t, False)
                    _set_edge(succ, pred, i, t, True)


def _fix_softmax_edges(nodes, succ, pred):
    """graph.py:553-574: edges around softmax (msa) nodes, consistent with DeepNets-1M graphs."""
    if not any('softmax' in nd.name.lower() for nd in nodes):
        return
    G = nx.DiGraph()
    G.add_nodes_from(range(len(nodes)))
    for a in range(len(nodes)):
        for b in sorted(succ[a]):
            G.add_edge(a, b)
    for i, nd in enumerate(nodes):
        if 'softmax' not in nd.name.lower():
            continue
        for out in sorted(succ[i]):
            for j in sorted(pred[out] - {i}):
                n_paths = 0
                for _ in nx.all_simple_paths(G, j, out):
                    n_paths += 1
                    if n_paths > 1:
                        break
                a_ij = j in succ[i]
                if n_paths > 1 or not a_ij:
                    _set_edge(succ, pred, j, out, False)
                if n_paths == 1 and not a_ij:
                    _set_edge(succ, pred, j, i, True)


def _fix_swin(nodes, succ, pred):
    """graph.py:579-598."""
    for i, nd in enumerate(nodes):
        low = nd.name.lower()
        if low.endswith('norm.weight'):
            for out in sorted(succ[i]):
                if nodes[out].name.endswith('norm1.weight') or 'Add' in nodes[out].name:
                    _set_edge(succ, pred, i, out, False)
                    target = nd.name.replace('norm', 'reduction')
                    for j, nd2 in enumerate(nodes):
                        if target in nd2.name:
                            _set_edge(succ, pred, i, j, True)
                            break
        elif low.endswith('attn.proj.bias'):
            for out in sorted(succ[i]):
                if nodes[out].name.endswith('reduction.weight'):
                    _set_edge(succ, pred, i, out, False)
                    for out2 in sorted(succ[out]):
                        if nodes[out2].name.startswith('AddBackward'):
                            _set_edge(succ, pred, i, out2, True)


# ----------------------------------------------------------------------------------------------------------------
def trace_model(model, list_all_nodes=False, reduce_graph=True, fix_weight_edges=True, fix_softmax_edges=True,
                verbose=True):
    """Returns (op ids [N], 1-hop edges [(src, dst)], node_info per cell, n_cells) of an nn.Module."""
    sz = getattr(model, 'expected_input_sz', 299 if isinstance(model, tvm.Inception3) else 224)
    input_sz = tuple(sz) if isinstance(sz, (tuple, list)) else (3, sz, sz)
    n_cells = getattr(model, '_n_cells', 1)
    table = _module_table()

    nodes, succ, pred = _autograd_graph(model, input_sz)
    if reduce_graph:
        nodes, succ, pred = _filter(nodes, succ, pred, table)
    if fix_weight_edges:
        _fix_weight_edges(nodes, succ, pred)
    if fix_softmax_edges:
        _fix_softmax_edges(nodes, succ, pred)
    if verbose and any(i in succ[i] for i in range(len(nodes))):
        print('WARNING: diagonal elements of the adjacency matrix should be zero')
    if isinstance(model, tvm.SwinTransformer):
        _fix_swin(nodes, succ, pred)
    if reduce_graph:
        nodes, succ, pred = _filter(nodes, succ, pred, table, only=['Add', 'Cat'])

    # input node: feeds every source node that is a weight (graph.py:604-613)
    n = len(nodes)
    nodes.append(_Node('input', None, None, None))
    succ.append(set())
    pred.append(set())
    for i in range(n + 1):
        if not pred[i] and 'weight' in nodes[i].name:
            _set_edge(succ, pred, n, i, True)
    n += 1
    for i in range(n):
        succ[i].discard(i)
        pred[i].discard(i)
    A = np.zeros((n, n), dtype=np.int64)
    for a in range(n):
        for b in succ[a]:
            A[a, b] = 1
    try:
        order = np.array(list(nx.topological_sort(nx.DiGraph(A))))
        nodes = [nodes[i] for i in order]
        A = A[order, :][:, order]
    except Exception as e:
        if verbose:
            print('WARNING: topological sort failed:', e)

    if isinstance(model, tvm.VisionTransformer):
        # the positional encoding gets an explicit sum node (graph.py:626-634, including its row/column shift)
        i = 0
        while i < len(nodes):
            if isinstance(nodes[i].module, tvm.vision_transformer.Encoder):
                nodes.insert(i + 1, _Node('AddBackward0', None, None, None))
                A = np.insert(A, i, 0, axis=0)
                A = np.insert(A, i, 0, axis=1)
                A[i, i + 1] = 1
            i += 1
    elif isinstance(model, tvm.SqueezeNet):
        assert nodes[-1].name.startswith('MeanBackward') and nodes[-3].name.startswith('classifier'), \
            (nodes[-1].name, nodes[-3].name)
        nodes.insert(len(nodes) - 3, copy.copy(nodes[-1]))
        del nodes[-1]

    # node features / node_info (graph.py:800-908)
    n = len(nodes)
    ops = np.empty(n, dtype=np.int64)
    info = [[] for _ in range(n_cells)]
    cell = 0
    for idx, nd in enumerate(nodes):
        pname = nd.name
        ci = _cell_index(pname, n_cells)
        if ci is not None:
            cell = ci
        p_stem, p_pos = pname.find('stem'), pname.find('pos_enc')
        if p_stem >= 0:
            pname = pname[p_stem:]
        elif p_pos >= 0:
            pname = pname[p_pos:]
        if nd.module is not None:
            parts = pname.split('.')
            for k, part in enumerate(parts):
                if part == '_ops' and k + 2 < len(parts) and parts[k + 2] != 'op':
                    try:
                        int(parts[k + 2])
                    except ValueError:
                        continue
                    parts.insert(k + 2, 'op')
                    pname = '.'.join(parts)
                    break
            prim = table[type(nd.module)](nd.module, pname)
        else:
            cut = pname.find('Backward')
            prim = _OP_TABLE.get(pname if cut == -1 else pname[:cut], 'sum')
            if n_cells > 1 and pname.startswith(('MaxPool', 'AvgPool')):
                pname = 'cells.%d.%s' % (cell, prim)
        size = None
        if nd.size is not None:
            size = tuple(nd.size)
        elif nd.is_op and nd.name != 'input' and 'pool' in prim:
            if nd.ksize is not None:
                size = (1, 1) + tuple(int(v.strip('() ')) for v in nd.ksize.split(','))
            else:
                size = (1, 1, 3, 3)
        if size is not None:
            if len(size) == 3 and size[0] == 1 and min(size[1:]) > 1:
                s = int(np.floor(size[1] ** 0.5))
                size = (1, size[2], s, s)
            elif len(size) == 4 and idx == n - 2 and max(size[2:]) == 1:
                size = size[:2]
        if prim not in _PRIM_ID:
            raise KeyError('Op/layer %s is not one of the DeepNets-1M primitives %s' % (prim, PRIMITIVES_DEEPNETS1M))
        ops[idx] = _PRIM_ID[prim]
        if nd.module is not None or 'pool' in prim or list_all_nodes:
            info[cell].append([idx, pname if nd.module is not None else prim, prim, size,
                               idx == n - 2 and '.weight' in pname, idx == n - 1 and '.bias' in pname])
    edges = np.argwhere(A == 1).astype(np.int32)
    return ops, edges, info, n_cells
