"""
Containers for computational graphs of neural networks: `Graph` and `GraphBatch`, call-compatible with the
reference's ghn3/graph.py (Graph: graph.py:282-352, GraphBatch: graph.py:38-88,155-185,243-279).

B200-first differences (the public attributes stay the same):
  * a Graph stores the 1-hop edge list; the shortest-path "virtual edge" matrix (`_Adj`, graph.py:755-798) is
    produced on the GPU by the bitset-BFS kernel (ghn3_spd_bfs) when the batch is moved to the device, instead of
    networkx on the host. Graph objects that already carry a dense `_Adj` (e.g. built by the reference, or loaded
    from DeepNets-1M with precomputed distances) are accepted and uploaded as they are;
  * a GraphBatch is varlen-packed on the device (no zero padding, no (B,N,N) int64 tensor, no mask): uint8 SPD
    matrices + uint16 (A_ij, A_ji) pair indices, see ghn3_b200/ops.py:GraphPack.

Tracing an nn.Module into a graph (reference graph.py:392-753) stays on the host: ghn3_b200/tracer.py.
"""
import numpy as np
import torch

from .weights import N_PRIMITIVES

PRIMITIVES_DEEPNETS1M = ['max_pool', 'avg_pool', 'sep_conv', 'dil_conv', 'conv', 'msa', 'cse', 'sum', 'concat',
                         'input', 'bias', 'bn', 'ln', 'pos_enc', 'glob_avg']
assert len(PRIMITIVES_DEEPNETS1M) == N_PRIMITIVES


class Graph:
    r"""
    Container for a computational graph of a neural network.

        graph = Graph(torchvision.models.resnet50())

    Either `model` or (`node_feat`, `node_info`, `A` | `edges`) must be given, as in the reference.
    """

    def __init__(self, model=None, node_feat=None, node_info=None, A=None, edges=None, net_args=None, net_idx=None,
                 ve_cutoff=50, list_all_nodes=False, reduce_graph=True, fix_weight_edges=True,
                 fix_softmax_edges=True, dense=False, verbose=True):
        assert node_feat is None or model is None, 'either model or other arguments must be specified'
        self.model = model
        self.ve_cutoff = ve_cutoff
        self._spd = None          # user-supplied dense distance matrix (numpy) if any
        self._edges1 = None       # (E, 2) int32 array of 1-hop edges
        self._verbose = verbose
        if model is not None:
            from .tracer import trace_model
            ops, edges1, info, n_cells = trace_model(model, list_all_nodes=list_all_nodes, reduce_graph=reduce_graph,
                                                     fix_weight_edges=fix_weight_edges,
                                                     fix_softmax_edges=fix_softmax_edges, verbose=verbose)
            self.n_cells = n_cells
            self.n_nodes = len(ops)
            self.node_feat = torch.as_tensor(np.asarray(ops, dtype=np.int64)).view(-1, 1)
            self.node_info = info
            self._edges1 = np.asarray(edges1, dtype=np.int32).reshape(-1, 2)
        else:
            self.n_nodes = len(node_feat)
            self.node_feat = torch.as_tensor(node_feat, dtype=torch.long).view(self.n_nodes, -1)
            self.node_info = node_info
            if A is not None:
                A = A.detach().cpu().numpy() if isinstance(A, torch.Tensor) else np.asarray(A)
                if (A > 1).any():
                    self._spd = A.astype(np.int64)               # already contains virtual edges
                self._edges1 = np.argwhere(A == 1).astype(np.int32)
            elif edges is not None:
                e = edges.detach().cpu().numpy() if isinstance(edges, torch.Tensor) else np.asarray(edges)
                if e.shape[1] >= 3:
                    if (e[:, 2] > 1).any():
                        spd = np.zeros((self.n_nodes, self.n_nodes), dtype=np.int64)
                        spd[e[:, 0], e[:, 1]] = e[:, 2]
                        self._spd = spd
                    e = e[e[:, 2] == 1]
                self._edges1 = e[:, :2].astype(np.int32)
            else:
                raise ValueError('Graph needs a model, an adjacency matrix A or an edge list')
        self.net_args = net_args
        self.net_idx = net_idx

    @classmethod
    def from_record(cls, rec, ve_cutoff=50):
        """Builds a Graph from a fixture record (tests/golden/graphs_tv.json.gz): ops, 1-hop edges, node_info."""
        g = cls.__new__(cls)
        g.model = None
        g.ve_cutoff = ve_cutoff
        g._spd = None
        g._verbose = False
        g.n_nodes = int(rec['n'])
        g.node_feat = torch.as_tensor(np.asarray(rec['ops'], dtype=np.int64)).view(-1, 1)
        g.node_info = [[[r[0], r[1], r[2], None if r[3] is None else tuple(r[3]), bool(r[4]), bool(r[5])]
                        for r in cell] for cell in rec['node_info']]
        g._edges1 = np.asarray(rec['edges'], dtype=np.int32).reshape(-1, 2)
        g.net_args = None
        g.net_idx = None
        g.n_cells = len(g.node_info)
        return g

    @property
    def edges1(self):
        return self._edges1

    @property
    def _Adj(self):
        """Dense (N, N) int64 matrix of shortest path distances (reference graph.py:798,904). Computed by the CUDA
        BFS kernel on first access when the graph was built from a model or a 1-hop edge list."""
        if self._spd is None:
            from . import ops
            if not torch.cuda.is_available():
                raise RuntimeError('ghn3_b200: Graph._Adj is produced by the CUDA shortest-path kernel; no CUDA '
                                   'device is available and there is no CPU fallback')
            pack = ops.GraphPack([self.n_nodes], edges=[self._edges1], cutoff=max(int(self.ve_cutoff), 1)).build()
            self._spd = pack.spd_matrix(0).cpu().numpy().astype(np.int64)
        return torch.as_tensor(self._spd, dtype=torch.long)

    @property
    def edges(self):
        """Sparse (E, 3) [row, col, distance] view of `_Adj` (reference graph.py:906-907)."""
        A = self._Adj
        ind = torch.nonzero(A)
        return torch.cat((ind, A[ind[:, 0], ind[:, 1]].view(-1, 1)), dim=1)

    def __len__(self):
        return self.n_nodes


_PACK_CACHE = {}      # (graph ids, cutoff, device) -> (graphs, GraphPack): host-side packing is done once per batch


def _norm_device(device):
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    return device


class GraphBatch:
    r"""
    Container for a batch of Graph objects (reference graph.py:38-88).

        batch = GraphBatch([Graph(torchvision.models.resnet50())], dense=True).to_device('cuda')
    """

    def __init__(self, graphs, dense=False):
        self.n_nodes, self.node_info, self.net_args, self.net_inds = [], [], [], []
        self.graphs = []
        self.dense = dense
        self.pack = None            # ops.GraphPack once on the device
        self.device = None
        if graphs is not None:
            if not isinstance(graphs, (list, tuple)):
                graphs = [graphs]
            for graph in graphs:
                self.append(graph)

    def append(self, graph):
        if self.pack is not None:
            raise RuntimeError('cannot append to a GraphBatch that is already on the device')
        graph = adopt_graph(graph)
        self.graphs.append(graph)
        self.n_nodes.append(graph.n_nodes)
        self.node_info.append(graph.node_info)
        self.net_args.append(graph.net_args)
        self.net_inds.append(graph.net_idx)
        if hasattr(graph, 'net'):
            if not hasattr(self, 'nets'):
                self.nets = []
            self.nets.append(graph.net)

    def to_device(self, device):
        if isinstance(device, (tuple, list)):
            device = device[0]
        device = _norm_device(device)
        if device.type != 'cuda':
            raise RuntimeError('ghn3_b200: GraphBatch can only be moved to a CUDA device (no CPU path)')
        if self.on_device(device):
            print('WARNING: GraphBatch is already on device %s.' % str(device))
            return self
        from . import ops
        cutoffs = {max(int(g.ve_cutoff), 1) for g in self.graphs}
        if len(cutoffs) > 1:
            raise RuntimeError('all graphs of a batch must use the same ve_cutoff')
        cutoff = cutoffs.pop() if cutoffs else 50
        key = (tuple(id(g) for g in self.graphs), cutoff, str(device))
        cached = _PACK_CACHE.get(key)
        if cached is not None and all(a is b for a, b in zip(cached[0], self.graphs)):
            # same (immutable) Graph objects as an earlier batch: reuse the packed host layout, repeat only the
            # H2D copy and the kernels
            self.pack = cached[1].clone_for_upload(device)
            self.pack.build()
            self.device = device
            return self
        op = np.concatenate([g.node_feat[:, 0].numpy() for g in self.graphs]).astype(np.int32)
        if all(g._spd is None for g in self.graphs):
            self.pack = ops.GraphPack(self.n_nodes, edges=[g.edges1 for g in self.graphs], cutoff=cutoff,
                                      device=device, op=op)
        else:
            vmax = max([cutoff] + [int(g._spd.max()) for g in self.graphs if g._spd is not None])
            self.pack = ops.GraphPack(self.n_nodes, spd=[g._Adj.numpy() for g in self.graphs], cutoff=vmax,
                                      device=device, op=op)
        self.pack.build()
        self.device = device
        if len(_PACK_CACHE) >= 1024:
            _PACK_CACHE.pop(next(iter(_PACK_CACHE)))
        _PACK_CACHE[key] = (list(self.graphs), self.pack)
        return self

    def on_device(self, device=None):
        if isinstance(device, (tuple, list)):
            device = device[0]
        return self.pack is not None and (device is None or _norm_device(device) == self.device)

    # reference-compatible tensor views (materialised on demand; the kernels do not use them)
    @property
    def node_feat(self):
        ops_ = torch.cat([g.node_feat[:, :1] for g in self.graphs]) if self.graphs else torch.zeros(0, 1).long()
        return ops_.to(self.device) if self.device is not None else ops_

    @property
    def edges(self):
        """(B, N, N) int64 zero-padded distance matrices, as in the reference's dense mode."""
        assert self.pack is not None, 'call to_device first'
        B, N = len(self.n_nodes), max(self.n_nodes)
        out = torch.zeros(B, N, N, dtype=torch.long, device=self.device)
        for b, n in enumerate(self.n_nodes):
            out[b, :n, :n] = self.pack.spd_matrix(b).long()
        return out

    @property
    def mask(self):
        B, N = len(self.n_nodes), max(self.n_nodes)
        m = torch.zeros(B, N, 1, dtype=torch.bool, device=self.device)
        for b, n in enumerate(self.n_nodes):
            m[b, :n] = True
        return m

    def to_dense(self, x=None):
        B, M, C = len(self.n_nodes), max(self.n_nodes), x.shape[-1]
        out = torch.zeros(B, M, C, device=x.device, dtype=x.dtype)
        offset = [0]
        for b in range(B):
            out[b, :self.n_nodes[b]] = x[offset[-1]: offset[-1] + self.n_nodes[b]]
            offset.append(offset[-1] + self.n_nodes[b])
        return out, offset

    def to_sparse(self, x):
        return torch.cat([x[b, :self.n_nodes[b]] for b in range(len(self.n_nodes))])

    def __getitem__(self, idx):
        return self.graphs[idx]

    def __len__(self):
        return len(self.n_nodes)

    def __iter__(self):
        for graph in self.graphs:
            yield graph


def adopt_graph(g):
    """Accepts our Graph, or any duck-compatible object (e.g. the reference's Graph): node_feat, node_info, _Adj."""
    if isinstance(g, Graph):
        return g
    if not (hasattr(g, 'node_feat') and hasattr(g, 'node_info')):
        raise TypeError('not a graph: %r' % type(g))
    A = getattr(g, '_Adj', None)
    new = Graph(node_feat=g.node_feat[:, :1] if torch.is_tensor(g.node_feat) else g.node_feat, node_info=g.node_info,
                A=A, edges=None if A is not None else getattr(g, 'edges', None),
                net_args=getattr(g, 'net_args', None), net_idx=getattr(g, 'net_idx', None))
    if hasattr(g, 'net'):
        new.net = g.net
    return new
