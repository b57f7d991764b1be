"""
Thin torch-tensor wrappers over the C ABI (include/ghn3_b200.h). PyTorch is used here only for device memory and
streams; every function enqueues our own kernels on the current CUDA stream and returns without synchronising.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from ._lib import BF16, TF32, F32, ACT_NONE, ACT_RELU, ACT_GELU  # noqa: F401

TORCH_DTYPE = {BF16: torch.bfloat16, TF32: torch.float32, F32: torch.float32}


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError('ghn3_b200: %s must be a CUDA tensor (there is no CPU path)' % name)


class HostBlob:
    """Packs several numpy arrays into ONE pinned host buffer -> ONE H2D copy; returns device views."""

    def __init__(self):
        self.parts = []
        self.size = 0

    def add(self, name, arr, align=16):
        arr = np.ascontiguousarray(arr)
        off = (self.size + align - 1) // align * align
        self.parts.append((name, off, arr))
        self.size = off + arr.nbytes
        return off

    def upload(self, device, pin=True):
        total = max(self.size, 16)
        host = getattr(self, 'host', None)
        if host is None:                      # packed once; repeated uploads copy from the same pinned buffer
            host = torch.empty(total, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
            hnp = host.numpy()
            for _, off, arr in self.parts:
                hnp[off:off + arr.nbytes] = arr.view(np.uint8).reshape(-1)
        dev = host.to(device, non_blocking=True)
        views = {}
        for name, off, arr in self.parts:
            tdt = {np.dtype('int32'): torch.int32, np.dtype('int64'): torch.int64, np.dtype('uint8'): torch.uint8,
                   np.dtype('float32'): torch.float32, np.dtype('int16'): torch.int16}[arr.dtype]
            views[name] = dev[off:off + arr.nbytes].view(tdt).reshape(arr.shape)
        self.host, self.dev, self.nbytes = host, dev, total
        return views


def derived_views(buf, layout):
    """(deg_in, deg_out, dist0, pair) views of a GraphPack.derived-style byte buffer."""
    n_deg, seg, n_pair = layout
    deg = [buf[i * seg:i * seg + n_deg * 4].view(torch.int32) for i in range(3)]
    return deg[0], deg[1], deg[2], buf[3 * seg:3 * seg + n_pair * 2].view(torch.int16)


class GraphPack:
    """
    Device-resident packed batch of graph structures: 1-hop edge lists (or user-supplied SPD matrices) on the way in;
    uint8 SPD, uint16 (A_ij, A_ji) pair index, degrees and input distance after `build()`.
    Layout per graph g: N_g x N_g matrices with leading dimension ld_g = round_up(N_g, 16) at offset mat_off[g].
    """

    def __init__(self, n_nodes, edges=None, spd=None, cutoff=50, device='cuda', op=None):
        self.device = torch.device(device)
        self.n_nodes = [int(n) for n in n_nodes]
        self.n_graphs = len(self.n_nodes)
        self.cutoff = int(cutoff)
        G = self.n_graphs
        self.node_off = np.zeros(G + 1, dtype=np.int32)
        self.node_off[1:] = np.cumsum(self.n_nodes)
        ld = [(n + 15) // 16 * 16 for n in self.n_nodes]
        self.ld = ld
        self.mat_off = np.zeros(G + 1, dtype=np.int64)
        self.mat_off[1:] = np.cumsum([n * l for n, l in zip(self.n_nodes, ld)])
        self.bits_off = np.zeros(G + 1, dtype=np.int64)
        self.bits_off[1:] = np.cumsum([n * ((n + 31) // 32) for n in self.n_nodes])
        self.total_nodes = int(self.node_off[-1])
        self.max_nodes = max(self.n_nodes) if G else 0
        blob = HostBlob()
        blob.add('node_off', self.node_off)
        blob.add('mat_off', self.mat_off)
        blob.add('bits_off', self.bits_off)
        self.has_edges = edges is not None
        if edges is not None:
            edges = [np.asarray(e, dtype=np.int32).reshape(-1, 2) for e in edges]
            self.edge_off = np.zeros(G + 1, dtype=np.int32)
            self.edge_off[1:] = np.cumsum([len(e) for e in edges])
            cat = np.concatenate(edges) if G else np.zeros((0, 2), np.int32)
            blob.add('edge_off', self.edge_off)
            blob.add('edge_src', np.ascontiguousarray(cat[:, 0]))
            blob.add('edge_dst', np.ascontiguousarray(cat[:, 1]))
            self.total_edges = int(self.edge_off[-1])
        else:
            assert spd is not None, 'either edges or spd matrices are required'
            packed = np.zeros(int(self.mat_off[-1]), dtype=np.uint8)
            for g, (n, l, A) in enumerate(zip(self.n_nodes, ld, spd)):
                A = np.asarray(A)
                if A.max(initial=0) > 254:
                    raise RuntimeError('ghn3_b200: shortest-path values above 254 are not supported')
                m = np.zeros((n, l), dtype=np.uint8)
                m[:, :n] = A
                packed[self.mat_off[g]:self.mat_off[g + 1]] = m.reshape(-1)
            blob.add('spd', packed)
        if op is not None:
            blob.add('op', np.asarray(op, dtype=np.int32))
        self.h2d_bytes = blob.size
        self._blob = blob
        self._upload()

    def _upload(self):
        v = self._blob.upload(self.device)
        self.d = v
        self.spd = v.get('spd')
        self.op_dev = v.get('op')
        self.pair = self.deg_in = self.deg_out = self.dist0 = self.derived = None

    def record_stream(self, stream):
        """Tells the caching allocator that `stream` reads this pack's buffers (they were allocated on another one)."""
        if getattr(self, '_recorded', None) is stream:
            return
        for t in (self._blob.dev, self.spd, getattr(self, 'derived', None), getattr(self, '_bits', None)):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(stream)
        self._recorded = stream

    def clone_for_upload(self, device=None):
        """A new pack with the same (immutable) host-side layout: only the H2D copy and the kernels are repeated."""
        import copy as _copy
        new = _copy.copy(self)
        if device is not None:
            new.device = torch.device(device)
        new._upload()
        return new

    def build(self, stream=None):
        """Runs the SPD BFS (if edges were given) and the derive kernel."""
        stream = L.current_stream() if stream is None else stream
        dev = self.device
        mat_total = int(self.mat_off[-1])
        if self.has_edges:
            self.spd = torch.empty(max(mat_total, 16), dtype=torch.uint8, device=dev)
            bits = torch.empty(max(int(self.bits_off[-1]), 4), dtype=torch.int32, device=dev)
            a = L.SpdArgs(n_graphs=self.n_graphs, cutoff=self.cutoff, node_off=L.ptr(self.d['node_off']),
                          edge_off=L.ptr(self.d['edge_off']), mat_off=L.ptr(self.d['mat_off']),
                          edge_src=L.ptr(self.d['edge_src']), edge_dst=L.ptr(self.d['edge_dst']),
                          max_nodes=self.max_nodes, total_nodes=self.total_nodes, total_edges=self.total_edges,
                          bits_total=int(self.bits_off[-1]), mat_total=mat_total, adj_bits=L.ptr(bits),
                          bits_off=L.ptr(self.d['bits_off']), spd=L.ptr(self.spd))
            L.call('spd_bfs', a, stream)
            self._bits = bits
        # one allocation for everything the derive kernel writes ([deg_in | deg_out | dist0 | pair], 256-byte aligned
        # parts): a program that replays a captured kernel sequence mirrors it with ONE device-to-device copy
        n_deg = max(self.total_nodes, 1)
        seg = (n_deg * 4 + 255) // 256 * 256
        self.derived = torch.empty(3 * seg + max(mat_total, 16) * 2, dtype=torch.uint8, device=dev)
        self.derived_layout = (n_deg, seg, max(mat_total, 16))
        self.deg_in, self.deg_out, self.dist0, self.pair = derived_views(self.derived, self.derived_layout)
        a = L.DeriveArgs(n_graphs=self.n_graphs, vmax=self.cutoff, node_off=L.ptr(self.d['node_off']),
                         mat_off=L.ptr(self.d['mat_off']), max_nodes=self.max_nodes, total_nodes=self.total_nodes,
                         spd=L.ptr(self.spd), pair=L.ptr(self.pair), deg_in=L.ptr(self.deg_in),
                         deg_out=L.ptr(self.deg_out), dist0=L.ptr(self.dist0))
        L.call('graph_derive', a, stream)
        return self

    def spd_matrix(self, g):
        """(N_g, N_g) uint8 tensor view (device) of graph g's SPD matrix."""
        n, l = self.n_nodes[g], self.ld[g]
        return self.spd[self.mat_off[g]:self.mat_off[g + 1]].view(n, l)[:, :n]

    def pair_matrix(self, g):
        n, l = self.n_nodes[g], self.ld[g]
        return self.pair[self.mat_off[g]:self.mat_off[g + 1]].view(n, l)[:, :n]


def node_features(op, shape_idx, pack, tables, hid):
    """tables: dict with embed_op, embed_ch, embed_sp, cent_in, cent_out, dist_embed (fp32 CUDA tensors)."""
    _require_cuda(op, 'op')
    x = torch.empty(op.numel(), hid, dtype=torch.float32, device=op.device)
    a = L.NodeFeaturesArgs(total_nodes=op.numel(), hid=hid, op=L.ptr(op), shape_idx=L.ptr(shape_idx),
                           deg_in=L.ptr(pack.deg_in), deg_out=L.ptr(pack.deg_out), dist0=L.ptr(pack.dist0),
                           embed_op=L.ptr(tables['embed_op']), embed_ch=L.ptr(tables['embed_ch']),
                           embed_sp=L.ptr(tables['embed_sp']), cent_in=L.ptr(tables['cent_in']),
                           cent_out=L.ptr(tables['cent_out']), dist_embed=L.ptr(tables['dist_embed']), x=L.ptr(x))
    L.call('node_features', a, L.current_stream())
    return x


def edge_lut(edge_embed, w1, b1, w2, b2, vmax=50, out=None):
    _require_cuda(edge_embed, 'edge_embed')
    C_, H = w1.shape[0], w2.shape[0]
    V = vmax + 1
    ws = torch.empty(2 * V * C_, dtype=torch.float32, device=w1.device)
    lut = torch.empty(H, V * V, dtype=torch.float32, device=w1.device) if out is None else out
    a = L.EdgeLutArgs(hid=C_, heads=H, vmax=vmax, edge_embed=L.ptr(edge_embed), w1=L.ptr(w1), b1=L.ptr(b1),
                      w2=L.ptr(w2), b2=L.ptr(b2), workspace=L.ptr(ws), lut=L.ptr(lut))
    L.call('edge_lut', a, L.current_stream())
    return lut


def layernorm(x, gamma, beta, out_dtype=BF16, dst_row=None, out_rows=None, out_f32=None):
    _require_cuda(x, 'x')
    rows, hid = x.shape
    out = torch.empty(rows if out_rows is None else out_rows, hid, dtype=TORCH_DTYPE[out_dtype], device=x.device)
    a = L.LayerNormArgs(rows=rows, hid=hid, x=L.ptr(x), gamma=L.ptr(gamma), beta=L.ptr(beta), out=L.ptr(out),
                        out_dtype=out_dtype, dst_row=L.ptr(dst_row), out_f32=L.ptr(out_f32))
    L.call('layernorm', a, L.current_stream())
    return out


def gemm(a, b, bias=None, act=ACT_NONE, in_dtype=BF16, out=None, out_dtype=F32, accumulate=False, problems=None,
         tiles=None, block_n=0, single=None, k_splits=0, x3=False, b_group=0, b_group_stride=0, bias_rows=False, b_dynamic=True, rowmap=None, swap_ab=False, persistent_single=False):
    """
    D = act(A @ B^T + bias). A [a_rows, K], B [b_rows, K] row-major CUDA tensors of the in_dtype storage type.
    Either a single problem covering all of A and B (default, or `single` = dict) or a grouped launch
    (`problems`, `tiles` device int32 tensors laid out as ghn3_gemm_problem / int32[4]).
    b_dynamic defaults to True here (safe for ad-hoc calls); the prediction program passes False for weights.
    """
    _require_cuda(a, 'a')
    K = a.shape[1]
    assert b.shape[1] == K and a.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        out = torch.empty(a.shape[0], b.shape[0], dtype=TORCH_DTYPE[out_dtype], device=a.device)
    g = L.GemmArgs(a=L.ptr(a), a_rows=a.shape[0], lda=a.stride(0), b=L.ptr(b), b_rows=b.shape[0], ldb=b.stride(0),
                   k=K, in_dtype=in_dtype, d=L.ptr(out), out_dtype=out_dtype, bias=L.ptr(bias), act=act,
                   accumulate=int(accumulate), block_n=block_n, k_splits=k_splits, tf32_x3=int(x3),
                   b_group=b_group, b_group_stride=b_group_stride, bias_rows=int(bias_rows),
                   b_dynamic=int(b_dynamic), rowmap=L.ptr(rowmap), swap_ab=int(swap_ab),
                   persistent_single=int(persistent_single))
    if problems is not None:
        g.problems = L.ptr(problems)
        g.tiles = L.ptr(tiles)
        g.n_tiles = tiles.shape[0]
    else:
        s = single or {}
        g.single = L.GemmProblem(a_row0=s.get('a_row0', 0), b_row0=s.get('b_row0', 0), m=s.get('m', a.shape[0]),
                                 n=s.get('n', b.shape[0]), d_off=s.get('d_off', 0),
                                 ldd=s.get('ldd', out.stride(0) if out.dim() == 2 else b.shape[0]),
                                 bias_off=s.get('bias_off', 0 if bias is not None else -1))
    L.call('gemm', g, L.current_stream())
    return out


def gemm_simt(a_ptr_tensor, sam, sak, b, sbn, sbk, bias, d, sdm, sdn, m, n, k, relu_a=False, act=ACT_NONE,
              batch=1, a_bs=0, d_bs=0, a_off=0, d_off=0):
    el = 4
    g = L.GemmSimtArgs(a=a_ptr_tensor.data_ptr() + a_off * el, sam=sam, sak=sak, b=L.ptr(b), sbn=sbn, sbk=sbk,
                       bias=L.ptr(bias), d=d.data_ptr() + d_off * el, sdm=sdm, sdn=sdn, m=m, n=n, k=k,
                       relu_a=int(relu_a), act=act, batch=batch, a_bs=a_bs, d_bs=d_bs)
    L.call('gemm_simt', g, L.current_stream())
    return d


def attention(qkv, pack, lut, hid, heads, dtype=BF16, out=None, lse2=None):
    """lse2 (optional, bf16 path): [heads, nodes] fp32 tensor that receives the log2-domain softmax statistics."""
    _require_cuda(qkv, 'qkv')
    if out is None:
        out = torch.empty(qkv.shape[0], hid, dtype=TORCH_DTYPE[dtype], device=qkv.device)
    a = L.AttentionArgs(n_graphs=pack.n_graphs, hid=hid, heads=heads, max_nodes=pack.max_nodes,
                        lut_size=lut.shape[1], node_off=L.ptr(pack.d['node_off']), mat_off=L.ptr(pack.d['mat_off']),
                        qkv=L.ptr(qkv), dtype=dtype, pair=L.ptr(pack.pair), lut=L.ptr(lut), out=L.ptr(out),
                        lse2=L.ptr(lse2), total_nodes=qkv.shape[0])
    L.call('attention', a, L.current_stream())
    return out


def scatter(descs_dev, n_descs, n_chunks, chunk_desc=None):
    a = L.ScatterArgs(descs=L.ptr(descs_dev), n_descs=n_descs, n_chunks=n_chunks, chunk_desc=L.ptr(chunk_desc))
    L.call('scatter', a, L.current_stream())


def convert(src, dst_dtype, out=None):
    _require_cuda(src, 'src')
    src = src.contiguous()
    dst = torch.empty(src.shape, dtype=TORCH_DTYPE[dst_dtype], device=src.device) if out is None else out
    assert dst.numel() == src.numel() and dst.is_contiguous()
    L.check(L.load().ghn3_convert_f32(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), C.c_int64(src.numel()),
                                      C.c_int32(dst_dtype), C.c_void_p(L.current_stream())), 'ghn3_convert_f32')
    return dst


# ----------------------------------------------------------------------------------------------------------------
# training path: thin wrappers of the adjoint kernels (see include/ghn3_b200.h, "Training path")
def _dt(t):
    return BF16 if t.dtype == torch.bfloat16 else F32


def transpose(src, dst_dtype=None, group=0, group_stride=0, rows=None, pad=8, out=None):
    """dst[c, r] = src[row(r), c]; the destination row stride is rounded up to `pad` elements (TMA alignment).
    Returns the [cols, rows] view of the padded buffer."""
    _require_cuda(src, 'src')
    assert src.dim() == 2 and src.stride(1) == 1
    rows = src.shape[0] if rows is None else rows
    cols = src.shape[1]
    dd = _dt(src) if dst_dtype is None else dst_dtype
    ld = (rows + pad - 1) // pad * pad
    if out is None:
        out = torch.zeros(cols, ld, dtype=TORCH_DTYPE[dd], device=src.device)
    a = L.TransposeArgs(src=L.ptr(src), src_dtype=_dt(src), ld_src=src.stride(0), rows=rows, cols=cols, group=group,
                        group_stride=group_stride, dst=L.ptr(out), dst_dtype=dd, ld_dst=out.stride(0))
    L.call('transpose', a, L.current_stream())
    return out[:, :rows]


def elementwise(op, a, b=None, out=None, out_dtype=None):
    _require_cuda(a, 'a')
    od = _dt(a) if out_dtype is None else out_dtype
    if out is None:
        out = torch.empty(a.shape, dtype=TORCH_DTYPE[od], device=a.device)
    args = L.ElementwiseArgs(op=op, n=a.numel(), a=L.ptr(a), a_dtype=_dt(a), b=L.ptr(b),
                             b_dtype=_dt(b) if b is not None else F32, out=L.ptr(out), out_dtype=od)
    L.call('elementwise', args, L.current_stream())
    return out


def colsum(src, dst, group=0, group_stride=0):
    _require_cuda(src, 'src')
    a = L.ColsumArgs(src=L.ptr(src), src_dtype=_dt(src), ld=src.stride(0), rows=src.shape[0], cols=src.shape[1],
                     group=group, group_stride=group_stride, dst=L.ptr(dst))
    L.call('colsum', a, L.current_stream())
    return dst


def layernorm_bwd(x, gamma, dy, dx, dgamma, dbeta, accumulate=False, dy_row=None):
    _require_cuda(x, 'x')
    a = L.LayerNormBwdArgs(rows=x.shape[0], hid=x.shape[1], x=L.ptr(x), gamma=L.ptr(gamma), dy=L.ptr(dy),
                           dy_dtype=_dt(dy), dy_row=L.ptr(dy_row), dx=L.ptr(dx), accumulate=int(accumulate),
                           dgamma=L.ptr(dgamma), dbeta=L.ptr(dbeta))
    L.call('layernorm_bwd', a, L.current_stream())
    return dx


def attention_bwd(qkv, out, d_out, pack, lut, hid, heads, d_lut=None, fwd_lse2=None):
    """fwd_lse2 (bf16 only): statistics kept by attention(..., lse2=...) -> tensor-core kernels; the edge-bias gradient
    then goes through a dS accumulation buffer and one binning pass (as the training program does for all layers)."""
    _require_cuda(qkv, 'qkv')
    M = qkv.shape[0]
    d_qkv = torch.empty_like(qkv)
    ws = torch.empty(2, heads, M, dtype=torch.float32, device=qkv.device)
    if fwd_lse2 is not None:
        ds_total = None
        if d_lut is not None:
            ds_total = torch.zeros(int(pack.mat_off[-1]) * heads, dtype=torch.float32, device=qkv.device)
        a = L.AttentionBwdArgs(n_graphs=pack.n_graphs, hid=hid, heads=heads, max_nodes=pack.max_nodes, total_nodes=M,
                               lut_size=lut.shape[1], node_off=L.ptr(pack.d['node_off']),
                               mat_off=L.ptr(pack.d['mat_off']), qkv=L.ptr(qkv), out=L.ptr(out), d_out=L.ptr(d_out),
                               dtype=_dt(qkv), pair=L.ptr(pack.pair), lut=L.ptr(lut), d_qkv=L.ptr(d_qkv), d_lut=None,
                               lse=L.ptr(ws[0]), delta=L.ptr(ws[1]), fwd_lse2=L.ptr(fwd_lse2), ds_total=L.ptr(ds_total))
        L.call('attention_bwd', a, L.current_stream())
        if d_lut is not None:
            b = L.LutBinArgs(n_graphs=pack.n_graphs, heads=heads, max_nodes=pack.max_nodes, lut_size=lut.shape[1],
                             node_off=L.ptr(pack.d['node_off']), mat_off=L.ptr(pack.d['mat_off']),
                             pair=L.ptr(pack.pair), ds_total=L.ptr(ds_total), d_lut=L.ptr(d_lut))
            L.call('lut_bin', b, L.current_stream())
        return d_qkv
    a = L.AttentionBwdArgs(n_graphs=pack.n_graphs, hid=hid, heads=heads, max_nodes=pack.max_nodes, total_nodes=M,
                           lut_size=lut.shape[1], node_off=L.ptr(pack.d['node_off']), mat_off=L.ptr(pack.d['mat_off']),
                           qkv=L.ptr(qkv), out=L.ptr(out), d_out=L.ptr(d_out), dtype=_dt(qkv), pair=L.ptr(pack.pair),
                           lut=L.ptr(lut), d_qkv=L.ptr(d_qkv), d_lut=L.ptr(d_lut), lse=L.ptr(ws[0]),
                           delta=L.ptr(ws[1]))
    L.call('attention_bwd', a, L.current_stream())
    return d_qkv


def edge_lut_bwd(edge_embed, w1, b1, w2, d_lut, vmax, grads):
    """grads: dict of zero-initialised fp32 tensors edge_embed, w1, b1, w2, b2 that receive the gradients (+=)."""
    C_, H = w1.shape[0], w2.shape[0]
    ws = torch.empty(4 * (vmax + 1) * C_, dtype=torch.float32, device=w1.device)
    a = L.EdgeLutBwdArgs(hid=C_, heads=H, vmax=vmax, edge_embed=L.ptr(edge_embed), w1=L.ptr(w1), b1=L.ptr(b1),
                         w2=L.ptr(w2), d_lut=L.ptr(d_lut), workspace=L.ptr(ws), d_edge_embed=L.ptr(grads['edge_embed']),
                         d_w1=L.ptr(grads['w1']), d_b1=L.ptr(grads['b1']), d_w2=L.ptr(grads['w2']),
                         d_b2=L.ptr(grads['b2']))
    L.call('edge_lut_bwd', a, L.current_stream())
    return grads


class FusedGraphormer:
    """Workspace + argument struct of ghn3_graphormer_fused (the persistent Graphormer-stack kernel, bf16) for one
    (stacked weights, total_nodes). `wstack` = dict of bf16 tensors w_qkv [L*3C, C], w_out [L*C, C], w_ff1 [L*4C, C],
    w_ff2 [L*C, 4C]; `layers_dev` = uint8 device tensor holding the ghn3_layer_weights table (fp32 vectors)."""

    def __init__(self, hid, heads, layers, wstack, layers_dev, total_nodes, device, x=None):
        N, C = total_nodes, hid
        self.wstack, self.layers_dev = wstack, layers_dev
        E = lambda *s, dtype=torch.bfloat16: torch.empty(*s, dtype=dtype, device=device)
        self.x = E(N, C, dtype=torch.float32) if x is None else x
        self.ao, self.qkv, self.ff = E(N, C), E(2, N, 3 * C), E(N, 4 * C)
        self.sync = torch.zeros(int(L.load().ghn3_graphormer_fused_sync_ints(N)), dtype=torch.int32, device=device)
        self.args = L.GraphormerFusedArgs(hid=C, heads=heads, layers=layers, w_qkv=L.ptr(wstack['w_qkv']),
                                          w_out=L.ptr(wstack['w_out']), w_ff1=L.ptr(wstack['w_ff1']),
                                          w_ff2=L.ptr(wstack['w_ff2']), layers_dev=L.ptr(layers_dev), total_nodes=N,
                                          x=L.ptr(self.x), ao=L.ptr(self.ao), qkv=L.ptr(self.qkv), ff=L.ptr(self.ff),
                                          sync=L.ptr(self.sync))

    def bind(self, pack, lut):
        a = self.args
        a.n_graphs, a.max_nodes, a.lut_size = pack.n_graphs, pack.max_nodes, lut.shape[1]
        a.node_off, a.mat_off = L.ptr(pack.d['node_off']), L.ptr(pack.d['mat_off'])
        a.pair, a.lut = L.ptr(pack.pair), L.ptr(lut)
        self._keep = (pack, lut)

    def run(self, stop_after=0, max_ctas=0):
        self.args.stop_after, self.args.max_ctas = stop_after, max_ctas
        L.call('graphormer_fused', self.args, L.current_stream())
        return self.x


def layer_table(layer_tensors, device):
    """Uploads a ghn3_layer_weights table: `layer_tensors` = list (per layer) of dicts name -> tensor."""
    import ctypes as ct
    tab = (L.LayerWeights * len(layer_tensors))()
    for l, t in enumerate(layer_tensors):
        for k, v in t.items():
            setattr(tab[l], k, v.data_ptr())
    raw = np.frombuffer(bytes(tab), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)
