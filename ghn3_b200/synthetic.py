"""
Synthetic large graphs for BASELINE.json config 4 ("thousands of nodes, stressing SPD-bias attention"):
a layered chain in which node i gets an edge from i-1 and, with probability 0.3, one from a uniform node in
[i-8, i-2]; primitive ids uniform over the 15 DeepNets-1M primitives; seed 0. torchvision has nothing beyond 816 nodes
(efficientnet_v2_l), so these DAGs are how the Graphormer stack is exercised at N = 1024 ... 4096.

`run_stack` drives node features -> SPD / edge indices -> L Graphormer layers -> final LayerNorm through the C ABI for a
graph that has no target network behind it (no decoders, no scatter).
"""
import numpy as np
import torch

from . import _lib as L
from . import ops


def synthetic_dag(n, seed=0, p_skip=0.3):
    """(edges [E, 2] int32 with local node ids, op ids [n] int32)."""
    rng = np.random.default_rng(seed)
    edges = [(i - 1, i) for i in range(1, n)]
    for i in range(10, n):
        if rng.random() < p_skip:
            edges.append((int(rng.integers(i - 8, i - 1)), i))
    op = rng.integers(0, 15, size=n).astype(np.int32)
    return np.asarray(edges, dtype=np.int32), op


class StackRunner:
    """Buffers + argument structs for the Graphormer stack of one packed batch of graphs without target networks."""

    def __init__(self, ghn, n_nodes_list, edges_list, op_list, device, shape_idx=None):
        self.ghn = ghn
        cfg = ghn.config
        C_, H = cfg['hid'], cfg['heads']
        self.pack = ops.GraphPack(list(n_nodes_list), edges=list(edges_list), cutoff=50 if ghn.ve else 1,
                                  device=device, op=np.concatenate(op_list).astype(np.int32))
        self.pack.build()
        w = ghn._device_weights()
        self.w = w
        N = self.pack.total_nodes
        if shape_idx is None:      # nodes without a parameter shape: the dummy rows of both tables
            n_ch, n_sp = w['tables']['embed_ch'].shape[0] - 1, w['tables']['embed_sp'].shape[0] - 1
            shape_idx = np.tile(np.array([n_ch, n_ch, n_sp, n_sp], dtype=np.int32), (N, 1))
        self.sidx = torch.from_numpy(np.ascontiguousarray(shape_idx, dtype=np.int32)).to(device)
        dt, x3 = w['dtype'], w['x3']
        tdt = ops.TORCH_DTYPE[w['act']]
        self.emb = torch.empty(N, C_, device=device)
        self.x = torch.empty(N, C_, device=device)
        self.h, self.qkv, self.ff = (torch.empty(N, k * C_, dtype=tdt, device=device) for k in (1, 3, 4))
        self.dec = torch.empty(N, C_, dtype=tdt, device=device)
        lut = ghn._lut(w, self.pack.cutoff)
        self.lut = lut
        t = w['tables']
        self.nf = L.NodeFeaturesArgs(total_nodes=N, hid=C_, op=L.ptr(self.pack.op_dev), shape_idx=L.ptr(self.sidx),
                                     deg_in=L.ptr(self.pack.deg_in), deg_out=L.ptr(self.pack.deg_out),
                                     dist0=L.ptr(self.pack.dist0), embed_op=L.ptr(t['embed_op']),
                                     embed_ch=L.ptr(t['embed_ch']), embed_sp=L.ptr(t['embed_sp']),
                                     cent_in=L.ptr(t['cent_in']), cent_out=L.ptr(t['cent_out']),
                                     dist_embed=L.ptr(t['dist_embed']), x=L.ptr(self.x))
        self.ga = L.GraphormerArgs(hid=C_, heads=H, layers=cfg['layers'], dtype=dt, layers_host=w['layers'],
                                   ln_w=L.ptr(w['ln_w']), ln_b=L.ptr(w['ln_b']), n_graphs=self.pack.n_graphs,
                                   total_nodes=N, max_nodes=self.pack.max_nodes, lut_size=lut.shape[1],
                                   node_off=L.ptr(self.pack.d['node_off']), mat_off=L.ptr(self.pack.d['mat_off']),
                                   pair=L.ptr(self.pack.pair), lut=L.ptr(lut), x=L.ptr(self.x), h=L.ptr(self.h),
                                   qkv=L.ptr(self.qkv), ff=L.ptr(self.ff), dec_in=L.ptr(self.dec), dec_dtype=w['act'],
                                   emb_f32=L.ptr(self.emb), tf32_x3=int(x3))

    def run(self):
        """Node features + the stack + the final LayerNorm; returns the fp32 node embeddings [total_nodes, C]."""
        st = L.current_stream()
        L.call('node_features', self.nf, st)
        L.call('graphormer_stack', self.ga, st)
        return self.emb

    def flops(self):
        """Algorithmic FLOPs of the stack: L * (24 * sum N * C^2 + 4 * sum N^2 * C)."""
        cfg = self.ghn.config
        C_, Ln = cfg['hid'], cfg['layers']
        ns = np.diff(self.pack.node_off.astype(np.int64)) if hasattr(self.pack, 'node_off') else None
        if ns is None:
            ns = np.array([self.pack.total_nodes])
        return float(Ln * (24.0 * ns.sum() * C_ * C_ + 4.0 * (ns.astype(np.float64) ** 2).sum() * C_))
