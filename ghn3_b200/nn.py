"""
GHN-3 on B200: the reference's Python API (`from_pretrained`, `GHN3(...)`, `ghn(model)`; reference ghn3/nn.py:31-349)
and state_dict layout over the C-ABI CUDA library. The modules below only HOLD parameters (same names and shapes
as the reference, so checkpoints load unchanged); all arithmetic of the forward pass runs in our kernels:

  graphs  --H2D-->  ghn3_spd_bfs / ghn3_graph_derive  ->  ghn3_node_features  ->  ghn3_graphormer_stack
          ->  decoder GEMMs (ghn3_gemm grouped: fc, conv.0, conv.2 restricted to the needed weight rows)
          ->  ghn3_gemm_simt (classification heads)  ->  ghn3_scatter straight into the target parameters

There is no CPU path: the GHN must live on a CUDA device.
"""
import ctypes as ct
import math
import weakref
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .graph import Graph, GraphBatch
from .plan import (BatchPlan, ModelPlan, SCATTER_CHUNK, SRC_CLSB, SRC_CLSW, SRC_D1, SRC_TOK, SRC_WOUT)
from .weights import (EDGE_EMBED_ROWS, MAX_DEGREE, MAX_INPUT_DIST, N_PRIMITIVES, channel_bins, normalize_config,
                      sinusoid_table, spatial_bins)

# compute_dtype -> (GEMM input storage dtype, 3-term compensated tf32?)
#   'bf16'   : bf16 operands, kind::f16 MMAs                                   (fast path, <= 2e-2)
#   'tf32'   : fp32 operands, kind::tf32 MMAs with hi/lo error compensation     (accurate path, <= 1e-3)
#   'tf32x1' : fp32 operands pre-rounded to tf32, single-pass kind::tf32 MMAs   (~1e-3, borderline)
DTYPES = {'bf16': (ops.BF16, False), 'tf32': (ops.TF32, True), 'tf32x1': (ops.TF32, False)}


# LayerNorm fused into the residual GEMMs (the CTA that completes a 128-row block normalises it). Implemented and
# tested (tests/test_kernels_gpu.py::test_gemm_fused_layernorm) but OFF: measured 2.9 ms vs 1.03 ms for the Graphormer
# stack on B200 -- one CTA normalising 128 x C values serially sits on the critical path, whereas the separate
# LayerNorm launch (58 CTAs, ~3 us with programmatic dependent launch) does not.
FUSE_LN = bool(int(__import__('os').environ.get('GHN3_FUSE_LN', '0')))

# The Graphormer stack as ONE persistent kernel (ghn3_graphormer_fused, csrc/graphormer_fused.cu) for bf16 predictions of
# small batches. Parity-tested (tests/test_fused_gpu.py, test_predict_gpu.py) and on par with the 7-launches-per-layer
# chain in isolation (1.04 vs 1.03 ms at 457 nodes, profiles/r2_fused_graphormer_trace.txt), but it holds every SM
# (one 205 KB CTA each), so the decoders of the previous call cannot run beside it: OFF by default, `ghn.fused_graphormer
# = True` or GHN3_FUSED=1 turns it on.
FUSED_DEFAULT = bool(int(__import__('os').environ.get('GHN3_FUSED', '0')))
FUSED_MAX_NODES = int(__import__('os').environ.get('GHN3_FUSED_MAX_NODES', '1024'))
# Inference programs replay their kernel sequence as CUDA graphs (ghn3_sequence_capture): one driver call instead of
# ~180 launches per prediction. `ghn.cuda_graphs = False` or GHN3_CUDA_GRAPHS=0 issues the kernels one by one.
GRAPHS_DEFAULT = bool(int(__import__('os').environ.get('GHN3_CUDA_GRAPHS', '1')))
# Two-lane decoder sequences: the 1-D decoder and the conv.2 column classes with few tiles run beside the large
# weight-streaming launches (GHN3_OP_FORK / JOIN); `ghn.decoder_lanes = False` or GHN3_LANES=0 keeps one lane.
LANES_DEFAULT = bool(int(__import__('os').environ.get('GHN3_LANES', '1')))
SCATTER_STREAM_DEFAULT = bool(int(__import__('os').environ.get('GHN3_SCATTER_STREAM', '1')))


def log(*a, **k):
    print(*a, **k)


# ----------------------------------------------------------------------------------------------------------------
# parameter containers with the reference's attribute names (state_dict contract)
# ----------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('parameter container: the computation runs in ghn3_b200 CUDA kernels')


class ShapeEncoder(_Holder):
    """ppuda ShapeEncoder parameters: embed_spatial [n_s+1, C/4], embed_channel [n_ch+1, C/4]."""

    def __init__(self, hid, num_classes, max_shape):
        super().__init__()
        self.embed_spatial = nn.Embedding(len(spatial_bins(max_shape)) + 1, hid // 4)
        self.embed_channel = nn.Embedding(len(channel_bins(num_classes)) + 1, hid // 4)


class EdgeEmbedding(_Holder):
    def __init__(self, hid, max_len):
        super().__init__()
        self.embed = nn.Embedding(max_len, hid)
        self.embed.weight.data = sinusoid_table(max_len, hid)      # graphormer.py:55-65


class Attention(_Holder):
    def __init__(self, dim, heads, edge_dim):
        super().__init__()
        self.num_heads = heads
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)                         # graphormer.py:89
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Identity())           # graphormer.py:92
        if edge_dim > 0:
            self.edge_embed = EdgeEmbedding(dim, EDGE_EMBED_ROWS)                 # graphormer.py:96
            self.proj_e = nn.Sequential(nn.Linear(edge_dim * dim, dim), nn.ReLU(), nn.Linear(dim, heads))


class FeedForward(_Holder):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Identity(), nn.Linear(hidden, dim),
                                 nn.Identity())                                    # graphormer.py:38-44


class GraphormerLayer(_Holder):
    def __init__(self, dim, num_heads, mlp_ratio=4, edge_dim=0, eps=1e-5, return_edges=False, **_):
        super().__init__()
        self.edge_dim = edge_dim
        self.return_edges = return_edges
        if edge_dim > 0:
            self.max_degree = MAX_DEGREE
            self.max_input_dist = MAX_INPUT_DIST
        self.ln1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, num_heads, edge_dim)
        self.ln2 = nn.LayerNorm(dim, eps=eps)
        self.ff = FeedForward(dim, int(dim * mlp_ratio))


class SequentialMultipleInOut(nn.Sequential):
    pass


class MLP(_Holder):
    def __init__(self, in_features, hid):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(in_features, hid[0]), nn.ReLU(), nn.Linear(hid[0], hid[1]), nn.Identity())


class ConvDecoder3(_Holder):
    """Parameters of the reference's ConvDecoder3 (nn.py:716-733): fc.0, conv.0, conv.2, class_layer_predictor.1."""

    def __init__(self, in_features, hid, out_shape, num_classes, is_ghn2=False):
        super().__init__()
        self.out_shape = out_shape
        self.num_classes = num_classes
        self.fc = nn.Sequential(nn.Linear(in_features, hid[0] * out_shape[2] * out_shape[3]), nn.ReLU())
        self.conv = nn.Sequential(nn.Linear(hid[0], hid[1]), nn.ReLU(),
                                  nn.Linear(hid[1], out_shape[0] * out_shape[1]), nn.Identity())
        self.class_layer_predictor = nn.Sequential(nn.ReLU(), nn.Linear(out_shape[0], num_classes))


class GHN(nn.Module):
    """Base-class marker (the reference's Trainer detects a GHN by isinstance(model, ppuda GHN), trainer.py:134)."""


class GHN3(GHN):
    r"""
    Transformer-based Graph HyperNetwork (GHN-3), B200-native. Constructor and forward signatures follow the
    reference (ghn3/nn.py:140-193). Extra keyword: compute_dtype='bf16' | 'tf32' selects the tensor-core path.
    """

    def __init__(self, max_shape, num_classes, hid, heads=8, layers=3, is_ghn2=False, pretrained=False, **kwargs):
        super().__init__()
        if is_ghn2:
            raise NotImplementedError('ghn3_b200 implements the GHN-3 (Graphormer) path only; GHN-2 checkpoints '
                                      '(is_ghn2=True, GatedGNN on sparse edge lists) are out of scope')
        kwargs.pop('act_layer', None)
        kwargs.pop('hypernet', None)
        kwargs.pop('decoder', None)
        self.weight_norm = kwargs.pop('weight_norm', False)
        self.ve = kwargs.pop('ve', False)
        self.layernorm = kwargs.pop('layernorm', False)
        self.debug_level = kwargs.pop('debug_level', 0)
        self.compute_dtype = kwargs.pop('compute_dtype', 'bf16')
        assert self.compute_dtype in DTYPES, self.compute_dtype
        if kwargs:
            raise TypeError('unexpected GHN3 arguments: %s' % sorted(kwargs))
        cfg = normalize_config(dict(max_shape=max_shape, num_classes=num_classes, hid=hid, heads=heads, layers=layers,
                                    layernorm=self.layernorm))
        self.config = cfg
        self.max_shape = cfg['max_shape']
        self.num_classes = num_classes
        self.hid, self.heads, self.layers = hid, heads, layers
        self._is_ghn2 = False
        ms = self.max_shape
        if self.layernorm:
            self.ln = nn.LayerNorm(hid)
        self.embed = nn.Embedding(N_PRIMITIVES, hid)
        self.shape_enc = ShapeEncoder(hid, num_classes, ms)
        self.gnn = SequentialMultipleInOut(*[
            GraphormerLayer(dim=hid, num_heads=heads, mlp_ratio=4, edge_dim=2 if layer == 0 else 0,
                            return_edges=layer < layers - 1) for layer in range(layers)])
        self.centrality_embed_in = nn.Embedding(MAX_DEGREE + 1, hid)
        self.centrality_embed_out = nn.Embedding(MAX_DEGREE + 1, hid)
        self.input_dist_embed = nn.Embedding(MAX_INPUT_DIST + 1, hid)
        self.decoder = ConvDecoder3(hid, (hid * 4, hid * 8), ms, num_classes)
        max_ch = max(ms[:2])
        self.decoder_1d = MLP(hid, (hid * 2, 2 * max_ch))
        self.bias_class = nn.Sequential(nn.ReLU(), nn.Linear(max_ch, num_classes))
        # nn.py:165-172
        for m in (self.decoder_1d.fc[-2], self.decoder.conv[-2], self.decoder.class_layer_predictor[-1]):
            m.weight.data /= 5.0
            m.bias.data *= 0
        for m in self.modules():                      # nn.py:170,704-713 (applies to every nn.Embedding)
            if isinstance(m, nn.Embedding):
                nn.init.trunc_normal_(m.weight.data, std=m.weight.shape[1] ** (-0.5))
        if not pretrained:
            self.fix_embed_layers()
        self._dev = None
        self._plan_cache = OrderedDict()

    # ------------------------------------------------------------------------------------------------------------
    def fix_embed_layers(self):
        """nn.py:174-184: the three structural embeddings live under gnn.0 in released checkpoints."""
        for name in ('centrality_embed_in', 'centrality_embed_out', 'input_dist_embed'):
            if name in self._modules:
                mod = self._modules.pop(name)
                setattr(self.gnn[0], name, mod)

    def is_dense(self):
        return True

    # ------------------------------------------------------------------------------------------------------------
    # device-side weight cache
    # ------------------------------------------------------------------------------------------------------------
    def _weights_signature(self):
        plist = self.__dict__.get('_param_list')
        if plist is None:
            plist = list(self.parameters())
            self.__dict__['_param_list'] = plist
        return tuple([p._version for p in plist]), self.compute_dtype

    def train(self, mode=True):
        # Tensor version counters do not see every in-place update (fused optimizers -- torch.optim.AdamW(fused=True)
        # -- leave them untouched), so switching mode, like every keep_grads forward, marks the device copies stale.
        self.__dict__['_dev_dirty'] = True
        return super().train(mode)

    def weights_updated(self):
        """Call after changing parameter values through a path that does not bump tensor versions."""
        self.__dict__['_dev_dirty'] = True

    def _apply(self, fn, *args, **kwargs):                 # .to() / .cuda() / .float(): parameters may be replaced
        self.__dict__['_param_list'] = None
        self._dev = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__['_param_list'] = None
        self._dev = None
        return super().load_state_dict(*args, **kwargs)

    def _device_weights(self):
        """Device copies of the weights in the compute dtype. Built once; when the parameters change IN PLACE
        (optimizer step, load_state_dict) the same buffers are refreshed, so prebuilt programs stay valid."""
        sig = self._weights_signature()
        dirty = self.__dict__.get('_dev_dirty', False)
        if self._dev is not None and self._dev['sig'] == sig and not dirty:
            return self._dev
        self.__dict__['_dev_dirty'] = False
        dev = self.embed.weight.device
        if dev.type != 'cuda':
            raise RuntimeError('ghn3_b200: the GHN must be on a CUDA device (got %s); there is no CPU path' % dev)
        self.fix_embed_layers()
        if self._dev is not None and self._dev['compute_dtype'] == self.compute_dtype and self._dev['ptrs'] == \
                [p.data_ptr() for p in self._param_list]:
            self._run_refresh(self._dev)
            self._dev['sig'] = sig
            return self._dev
        dt, x3 = DTYPES[self.compute_dtype]
        C, S = self.hid, self.max_shape[2]
        refresh = []

        def f(p):
            """fp32 view of a parameter: aliases its storage when possible, otherwise a copy that is refreshed."""
            d = p.detach()
            if d.dtype == torch.float32 and d.is_contiguous():
                return d
            buf = d.contiguous().float()
            refresh.append(lambda: buf.copy_(p.detach()))
            return buf

        refresh_ops = []                                   # (entry point, args): replayed by ONE ghn3_run_sequence call

        def cv(p, out=None):
            if x3:                                         # x3 keeps the fp32 master weights
                return f(p)
            src = f(p)
            buf = ops.convert(src, dt, out=out)
            if src.data_ptr() == p.data_ptr():             # aliasing fp32 parameter: a prebuilt conversion op
                refresh_ops.append(('elementwise', L.ElementwiseArgs(op=L.EW_COPY, n=src.numel(), a=src.data_ptr(),
                                                                     a_dtype=ops.F32, b=None, b_dtype=0,
                                                                     out=buf.data_ptr(), out_dtype=dt)))
            else:
                refresh.append(lambda: ops.convert(p.detach().float().contiguous(), dt, out=buf))
            return buf

        g0 = self.gnn[0]
        w = {'sig': sig, 'dtype': dt, 'x3': x3, 'act': ops.F32 if x3 else dt, 'compute_dtype': self.compute_dtype,
             'ptrs': [p.data_ptr() for p in self._param_list], 'refresh': refresh, 'refresh_ops': refresh_ops,
             'refresh_seq': None}
        w['tables'] = {'embed_op': f(self.embed.weight), 'embed_ch': f(self.shape_enc.embed_channel.weight),
                       'embed_sp': f(self.shape_enc.embed_spatial.weight),
                       'cent_in': f(g0.centrality_embed_in.weight), 'cent_out': f(g0.centrality_embed_out.weight),
                       'dist_embed': f(g0.input_dist_embed.weight)}
        self._lut_cache = {}
        w['lut_inputs'] = (f(g0.attn.edge_embed.embed.weight), f(g0.attn.proj_e[0].weight), f(g0.attn.proj_e[0].bias),
                           f(g0.attn.proj_e[2].weight), f(g0.attn.proj_e[2].bias))

        def refresh_luts():
            for vmax, lut in self._lut_cache.items():
                ops.edge_lut(*w['lut_inputs'], vmax=vmax, out=lut)
        refresh.append(refresh_luts)
        layers = (L.LayerWeights * self.layers)()
        keep = []
        # bf16 GEMM weights of all layers stacked along rows: four TMA descriptors cover the stack (fused kernel); the
        # per-layer tensors used everywhere else are views of the stacks
        stacks = None
        if not x3 and dt == ops.BF16:
            Ln = self.layers
            stacks = {'w_qkv': torch.empty(Ln * 3 * C, C, dtype=torch.bfloat16, device=dev),
                      'w_out': torch.empty(Ln * C, C, dtype=torch.bfloat16, device=dev),
                      'w_ff1': torch.empty(Ln * 4 * C, C, dtype=torch.bfloat16, device=dev),
                      'w_ff2': torch.empty(Ln * C, 4 * C, dtype=torch.bfloat16, device=dev)}

        def cvs(p, kind, l):
            if stacks is None:
                return cv(p)
            rows = p.shape[0]
            return cv(p, out=stacks[kind][l * rows:(l + 1) * rows])
        for l, layer in enumerate(self.gnn):
            t = dict(ln1_w=f(layer.ln1.weight), ln1_b=f(layer.ln1.bias), w_qkv=cvs(layer.attn.to_qkv.weight, 'w_qkv', l),
                     w_out=cvs(layer.attn.to_out[0].weight, 'w_out', l), b_out=f(layer.attn.to_out[0].bias),
                     ln2_w=f(layer.ln2.weight), ln2_b=f(layer.ln2.bias), w_ff1=cvs(layer.ff.net[0].weight, 'w_ff1', l),
                     b_ff1=f(layer.ff.net[0].bias), w_ff2=cvs(layer.ff.net[3].weight, 'w_ff2', l),
                     b_ff2=f(layer.ff.net[3].bias))
            keep.append(t)
            for k, v in t.items():
                setattr(layers[l], k, v.data_ptr())
        w['layers'], w['layers_keep'] = layers, keep
        w['wstack'] = stacks
        # the same table on the device (the fused kernel reads the fp32 vectors' addresses from it)
        w['layers_dev'] = torch.from_numpy(np.frombuffer(bytes(layers), dtype=np.uint8).copy()).to(dev)
        # layernorm=False (nn.py:262): no final LayerNorm -- the kernel runs in its identity form (NULL gamma / beta)
        w['ln_w'], w['ln_b'] = (f(self.ln.weight), f(self.ln.bias)) if self.layernorm else (None, None)
        dec = self.decoder
        # fc weight repacked position-major: [c*S*S + p][k] -> [p][c][k], so one decoder-grid position is one
        # contiguous [4C, C] block and a crop window is a set of row ranges (nn.py:738-745)
        repack_w = lambda: f(dec.fc[0].weight).view(4 * C, S * S, C).permute(1, 0, 2).contiguous().view(S * S * 4 * C, C)
        repack_b = lambda: f(dec.fc[0].bias).view(4 * C, S * S).t().contiguous().view(-1)
        if x3:
            w['fc_w'] = repack_w()
            refresh.append(lambda: w['fc_w'].copy_(repack_w()))
        else:
            w['fc_w'] = ops.convert(repack_w(), dt)
            refresh.append(lambda: ops.convert(repack_w(), dt, out=w['fc_w']))
        w['fc_b'] = repack_b()
        refresh.append(lambda: w['fc_b'].copy_(repack_b()))
        w['c0_w'], w['c0_b'] = cv(dec.conv[0].weight), f(dec.conv[0].bias)
        w['c2_w'], w['c2_b'] = cv(dec.conv[2].weight), f(dec.conv[2].bias)
        w['cls_w'], w['cls_b'] = cv(dec.class_layer_predictor[1].weight), f(dec.class_layer_predictor[1].bias)
        w['d1_w0'], w['d1_b0'] = cv(self.decoder_1d.fc[0].weight), f(self.decoder_1d.fc[0].bias)
        w['d1_w1'], w['d1_b1'] = cv(self.decoder_1d.fc[2].weight), f(self.decoder_1d.fc[2].bias)
        w['bc_w'], w['bc_b'] = f(self.bias_class[1].weight), f(self.bias_class[1].bias)
        self._dev = w
        self._compute_cache_io(w)
        return w

    _CACHED = ('fc_w', 'c0_w', 'c2_w', 'cls_w', 'd1_w0', 'd1_w1')

    def _compute_cache_io(self, w):
        """Reads (if valid) or writes the compute-dtype weight copies beside the checkpoint (from_pretrained(...,
        cache_compute_copy=True)). Validity = same checkpoint size / mtime, same dtype, same tensor shapes."""
        path = self.__dict__.get('_compute_cache_path')
        if path is None or w['x3']:
            return
        import os
        tensors = {k: w[k] for k in self._CACHED}
        if w['wstack'] is not None:
            tensors.update({'stack.' + k: v for k, v in w['wstack'].items()})
        key = list(self._compute_cache_key) + [self.compute_dtype] + [list(t.shape) for t in tensors.values()]
        if os.path.exists(path):
            try:
                blob = torch.load(path, map_location='cpu', weights_only=False)
                if blob.get('key') == key:
                    for k, t in tensors.items():
                        t.copy_(blob['tensors'][k])
                    self.__dict__['_compute_cache_hit'] = True
                    return
            except Exception:
                pass
        torch.save({'key': key, 'tensors': {k: t.cpu() for k, t in tensors.items()}}, path)
        self.__dict__['_compute_cache_hit'] = False

    _REFRESH_OPC = {'elementwise': 10, 'transpose': 9}

    def _run_refresh(self, w):
        """Re-derives every device copy from the current parameter values (same buffers, same addresses)."""
        self.flush_all()                         # overlapped predictions may still be reading the old copies
        ops_ = w['refresh_ops']
        if ops_:
            seq = w['refresh_seq']
            if seq is None or len(seq) != len(ops_):
                seq = (L.SeqOp * len(ops_))()
                for i, (name, args) in enumerate(ops_):
                    seq[i].op = self._REFRESH_OPC[name]
                    seq[i].args = ct.cast(ct.pointer(args), ct.c_void_p)
                w['refresh_seq'] = seq
            L.check(L.load().ghn3_run_sequence(seq, len(ops_), ct.c_void_p(L.current_stream())),
                    'ghn3_run_sequence (weight refresh)')
        for fn in w['refresh']:
            fn()

    def _lut(self, w, vmax):
        if vmax not in self._lut_cache:
            self._lut_cache[vmax] = ops.edge_lut(*w['lut_inputs'], vmax=vmax)
        return self._lut_cache[vmax]

    # ------------------------------------------------------------------------------------------------------------
    def forward(self, nets_torch, graphs=None, return_embeddings=False, predict_class_layers=True,
                bn_track_running_stats=True, keep_grads=False, reduce_graph=False):
        r"""
        Predict parameters for a list of >=1 networks (signature of the reference, nn.py:186-209).
        The predicted tensors are written in place into the parameters of `nets_torch` on the GHN's device.
        """
        device = self.embed.weight.device
        if device.type != 'cuda':
            raise RuntimeError('ghn3_b200: the GHN must be on a CUDA device (got %s); there is no CPU path' % device)
        keep_grads = self.training if keep_grads is None else bool(keep_grads)       # nn.py:525
        is_lst = isinstance(nets_torch, (list, tuple))
        nets = list(nets_torch) if is_lst else [nets_torch]
        if len(nets) == 0:                                   # empty batch: nothing to predict
            return ([], torch.empty(0, self.hid, device=device)) if return_embeddings else []

        if graphs is None:
            graphs = GraphBatch([self._traced_graph(net) for net in nets], dense=True)
        elif not isinstance(graphs, GraphBatch):
            graphs = GraphBatch(list(graphs) if isinstance(graphs, (list, tuple)) else [graphs], dense=True)
        if not graphs.on_device(device):
            graphs.to_device(device)
        assert len(graphs) == len(nets), 'number of graphs and networks must match'

        if keep_grads:                     # a training step: the optimizer has (probably) just moved the weights
            self.__dict__['_dev_dirty'] = True
        w = self._device_weights()
        bp = self._batch_plan(graphs, nets, predict_class_layers, reduce_graph)
        if keep_grads:
            self.__dict__['_dev_dirty'] = True
            from .train import forward_keep_grads
            emb = forward_keep_grads(self, nets, graphs, w, bp, return_embeddings)
        else:
            emb = self._run(w, graphs.pack, bp, return_embeddings)

        if bn_track_running_stats is None:
            bn_track_running_stats = self.training
        if not bn_track_running_stats:
            def bn_set_train(module):                      # nn.py:333-342
                if isinstance(module, nn.BatchNorm2d):
                    module.track_running_stats = False
                    module.training = True
            for net in nets:
                net.apply(bn_set_train)

        if self.debug_level:
            n_params = sum(sum(p.numel() for p in net.parameters()) for net in nets)
            n_pred = sum(p.n_params for p in bp.plans)
            log('number of parameter tensors predicted using GHN: {}, total parameters predicted: {} ({})'.format(
                sum(p.n_tensors for p in bp.plans), n_pred,
                'MATCHED!' if n_params == n_pred else 'ERROR! NOT MATCHED WITH {} ACTUAL PARAMS'.format(n_params)))

        out = nets if is_lst or len(nets) > 1 else nets[0]
        return (out, emb) if return_embeddings else out

    # ------------------------------------------------------------------------------------------------------------
    def _traced_graph(self, net):
        """Graph(net) with a per-model cache keyed by an architecture signature (module classes + parameter names and
        shapes): `ghn(model)` called again on the same architecture does not re-trace it."""
        cutoff = 50 if self.ve else 1
        sig = (cutoff, tuple((n, type(m).__name__) for n, m in net.named_modules()),
               tuple((n, tuple(p.shape)) for n, p in net.named_parameters()))
        hit = net.__dict__.get('_ghn3_b200_graph')
        if hit is not None and hit[0] == sig:
            return hit[1]
        g = Graph(net, ve_cutoff=cutoff)
        net.__dict__['_ghn3_b200_graph'] = (sig, g)
        return g

    def _batch_plan(self, graphs, nets, predict_class_layers, reduce_graph):
        plans = []
        for graph, net in zip(graphs.graphs, nets):
            cache = net.__dict__.setdefault('_ghn3_b200_plans', {})
            key = (id(graph), self.max_shape, self.num_classes, predict_class_layers)
            plan = cache.get(key)
            if plan is None or plan.graph_ref() is not graph:
                plan = ModelPlan(graph, net, self.config, predict_class_layers)
                plan.graph_ref = weakref.ref(graph)
                if reduce_graph:
                    plan.prune_unmatched()
                cache.clear()
                cache[key] = plan
            plans.append(plan)
        key = tuple(id(p) for p in plans)
        bp = self._plan_cache.get(key)
        if bp is None or any(a is not b for a, b in zip(bp.plans, plans)):
            bp = BatchPlan(plans, self.config)
            self._plan_cache[key] = bp
            # A batch plan owns its device programs (activation workspaces; in training the saved activations of every
            # layer, the prediction buffer and a full-size flat gradient buffer) and pins its target networks. Meta-batches
            # of real GHN training never repeat, so only the few most recent plans are kept: `plan_cache_size` (default 4
            # in training mode, 1024 plans for inference, where programs are small and loops over a fixed model zoo
            # do repeat).
            limit = getattr(self, 'plan_cache_size', None)
            if limit is None:
                limit = 4 if self.training else 1024
            while len(self._plan_cache) > max(1, int(limit)):
                _, old = self._plan_cache.popitem(last=False)
                # drop the programs' buffers explicitly: program -> pred_flat -> grad_fn -> ctx -> program is a cycle
                # through C++ autograd nodes that the garbage collector does not see
                progs = list(old.__dict__.get('programs', [])) + list(old.__dict__.get('programs_shared', [])) + \
                    [old.__dict__.get('program'),
                                                                   old.__dict__.get('train_program')]
                for attr in ('program', 'programs', 'programs_shared', 'train_program', 'dev', 'out_meta'):
                    old.__dict__.pop(attr, None)
                for prog in progs:
                    if prog is not None and prog is not self.__dict__.get('last_program'):
                        prog.__dict__.clear()
        else:
            self._plan_cache.move_to_end(key)
        return bp

    def _static_device(self, bp, device):
        """Uploads the per-plan static metadata once (op ids come from the graphs, see _run)."""
        st = getattr(bp, 'dev', None)
        if st is not None and st['device'] == device:
            return st
        blob = ops.HostBlob()
        blob.add('shape_idx', bp.shape_idx.astype(np.int32))
        blob.add('dst_row', bp.dst_row)
        blob.add('chunk_desc', bp.chunk_desc)
        blob.add('fc_problems', bp.fc_problems.view(np.uint8))
        blob.add('fc_tiles', bp.fc_tiles)
        blob.add('fc_tiles_swap', bp.fc_tiles_swap)
        blob.add('fc_rowmap', bp.fc_rowmap if len(bp.fc_rowmap) else np.zeros(1, np.int32))
        for g_, pr, tl in bp.c2_launches:
            blob.add('c2_%d_problems' % g_, pr.view(np.uint8))
            blob.add('c2_%d_tiles' % g_, tl)
        st = blob.upload(device)
        st['device'] = device
        st['bytes'] = blob.size
        bp.dev = st
        return st

    def _run(self, w, pack, bp, return_embeddings):
        """Executes the device program of this (batch plan, weight version): ONE C call enqueues every kernel."""
        device = self.embed.weight.device
        prof = getattr(self, '_profile', None)
        overlap = bool(getattr(self, 'overlap_scatter', False)) and prof is None
        # `pipeline_depth` (overlapped mode only): that many programs -- each with its own activation buffers and its own
        # high-priority stream -- take successive calls on the same batch plan in turn, so the latency-bound Graphormer
        # chains of calls k and k+1 run side by side; decoders + scatter of every call stay ordered on ONE side stream
        depth = max(1, int(getattr(self, 'pipeline_depth', 1))) if overlap else 1
        # programmatic dependent launch: lowest latency for one chain, lower throughput for >= 3 chains side by side
        # (each chain parks its next kernel's CTAs on the SMs); `programmatic_launch` = True / False overrides
        pdl = getattr(self, 'programmatic_launch', None)
        L.set_programmatic_launch(depth < 3 if pdl is None else pdl)
        # throughput mode: the persistent decoder GEMMs leave ~1/3 of the SMs to the other predictions' Graphormer chains
        # (`persistent_ctas` overrides; 0 = one CTA per SM)
        cap = getattr(self, 'persistent_ctas', None)
        if cap is None:
            cap = (2 * torch.cuda.get_device_properties(device).multi_processor_count + 2) // 3 if depth >= 3 else 0
        L.set_persistent_ctas(cap)
        share = not overlap and prof is None and bool(getattr(self, 'share_workspaces', True))
        progs = bp.__dict__.setdefault('programs_shared' if share else 'programs', [])
        want = bool(return_embeddings)
        if any(p_.w is not w or p_.device != device or p_.want_emb != want for p_ in progs):
            self.flush_all()
            del progs[:]
        slot = bp.__dict__.get('next_slot', 0) % depth
        while len(progs) <= slot:
            progs.append(_Program(self, w, bp, device, want, slot=len(progs), shared=share))
        prog = progs[slot]
        bp.next_slot = slot + 1
        bp.program = prog
        prog.bind_pack(pack)
        prog.refresh_targets(self.weight_norm)
        prog.run(prof, overlap=overlap)
        self.last_program = prog
        return prog.emb

    def _overlap_streams(self, device, slot):
        """(high-priority stream of program slot `slot`, the one side stream every program's decoders + scatter use)."""
        ov = self.__dict__.setdefault('_ov_streams', {})
        st = ov.get(device)
        if st is None:
            st = ov[device] = {'side': torch.cuda.Stream(device=device), 'hi': []}
            # the write-bound scatter gets its own stream so that it runs beside the NEXT call's read-bound decoder
            # GEMMs (HBM reads and writes together reach the copy bandwidth; either alone does not);
            # `ghn.scatter_stream = False` keeps it behind the decoders on one stream
            st['sc'] = torch.cuda.Stream(device=device) if getattr(self, 'scatter_stream', SCATTER_STREAM_DEFAULT) else st['side']
        while len(st['hi']) <= slot:
            st['hi'].append(torch.cuda.Stream(device=device, priority=-1))
        return st['hi'][slot], st['side'], st['sc']

    def flush_all(self):
        """Waits (on the current stream) for every overlapped prediction in flight, whatever program ran it."""
        for st in self.__dict__.get('_ov_streams', {}).values():
            for strm in [st['side'], st['sc']] + st['hi']:
                torch.cuda.current_stream().wait_stream(strm)

    def result_stream(self):
        """The stream on which the last prediction's parameters (and predicted_sumsq) become final: the side stream in
        `overlap_scatter` mode, else the current stream. Work that consumes them without stalling the next prediction
        (e.g. param_norms + a D2H copy) can be enqueued there: `with torch.cuda.stream(ghn.result_stream()): ...`."""
        prog = getattr(self, 'last_program', None)
        if prog is not None and getattr(prog, 'ev_scatter', None) is not None:
            return prog.sc
        return torch.cuda.current_stream()

    def flush(self):
        """With `overlap_scatter` the final tile/normalise/scatter kernel of a prediction runs on a side stream so that
        it overlaps the next prediction's Graphormer stack; flush() makes the current stream wait for it. Call it
        before the predicted parameters (or returned embeddings) are used or timed."""
        prog = getattr(self, 'last_program', None)
        ev = getattr(prog, 'ev_scatter', None) if prog is not None else None
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def param_norms(self, nets):
        """Total L2 norm of each network's parameters after the last forward call (the reference's norm_check metric,
        nn.py:783-797) as a float64 device tensor [n_models], without re-reading the predicted tensors: their sum of
        squares comes from the scatter kernel; only the (few, small) parameters the GHN does not predict are read."""
        prog = self.last_program
        nets = list(nets) if isinstance(nets, (list, tuple)) else [nets]
        cache = prog.__dict__.setdefault('_unpred', None)
        if cache is None:
            written = {(id(m), a) for (m, a, _, _) in prog.bp.desc_targets}
            cache = []
            for net in nets:
                rest = []
                for mod in net.modules():
                    for a, p in mod._parameters.items():
                        if p is not None and (id(mod), a) not in written:
                            rest.append(p)
                cache.append(rest)
            prog._unpred = cache
            prog._unpred_out = torch.zeros(len(nets), dtype=torch.float64, device=prog.device)
            prog._unpred_meta = [None] * len(nets)
        out = prog._unpred_out
        for i, rest in enumerate(cache):
            if not rest:
                continue
            ptrs = [p.data_ptr() for p in rest]
            meta = prog._unpred_meta[i]
            if meta is None or meta[0] != ptrs:
                arr = torch.from_numpy(np.array([ptrs, [p.numel() for p in rest]], dtype=np.int64)).to(prog.device)
                meta = (ptrs, arr, L.SumsqArgs(ptrs=arr[0].data_ptr(), numels=arr[1].data_ptr(), n=len(rest)))
                prog._unpred_meta[i] = meta
            meta[2].out = out[i:i + 1].data_ptr()
            L.call('sumsq', meta[2], L.current_stream())
        return (prog.pred_sumsq + out).sqrt_()

    def predicted_sumsq(self):
        """Device tensor [n_models] (float64): sum of squares of every parameter value written by the last forward
        call, accumulated inside the scatter kernel (no second pass over the parameters, no host sync)."""
        return self.last_program.pred_sumsq


class _SharedWorkspace:
    """Grow-only temporaries shared by the serial-mode programs of one (GHN, device, stream): the k-th buffer a program
    asks for is a typed view of the k-th shared block; a block that is too small is replaced by a larger one (programs
    built earlier keep their views of the old block alive)."""

    def __init__(self, device):
        self.device = device
        self.blocks = []
        self.k = 0

    @classmethod
    def of(cls, ghn, device):
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        table = ghn.__dict__.setdefault('_shared_ws', {})
        ws = table.get(key)
        if ws is None:
            ws = table[key] = cls(device)
        ws.k = 0                               # a new program starts at the first block
        return ws

    def empty(self, shape, dtype):
        numel = 1
        for v in shape:
            numel *= int(v)
        nbytes = max(numel * dtype.itemsize, 1)
        k = self.k
        self.k += 1
        if k == len(self.blocks):
            self.blocks.append(None)
        blk = self.blocks[k]
        if blk is None or blk.numel() < nbytes:
            # grow with some head-room so that a zoo of similar models settles after a few allocations
            blk = self.blocks[k] = torch.empty((nbytes * 5 // 4 + 255) // 256 * 256, dtype=torch.uint8,
                                               device=self.device)
        return blk[:numel * dtype.itemsize].view(dtype).view(*[int(v) for v in shape])


def _destroy_sequences(graphs):
    """Frees the captured kernel sequences (CUDA graph executables) of a program."""
    lib = L.load()
    while graphs:
        _, h = graphs.popitem()
        lib.ghn3_sequence_destroy(ct.c_void_p(h))


class _Program:
    """
    The kernel sequence of one batch plan with every argument struct prebuilt and every workspace buffer allocated
    once; per call only the graph-pack pointers and (if they moved) the target-parameter addresses are patched.
    """
    OP = {'node_features': 1, 'graphormer_stack': 2, 'gemm': 3, 'gemm_simt': 4, 'scatter': 5, 'relu_transpose': 6,
          'graphormer_train_fwd': 7, 'layernorm': 21, 'graphormer_fused': 22}

    def __init__(self, ghn, w, bp, device, want_emb, train=False, slot=0, shared=False):
        self.w, self.bp, self.device, self.want_emb = w, bp, device, want_emb
        self.train = train
        self.slot = slot
        dt, x3, act = w['dtype'], w['x3'], w['act']
        tdt = ops.TORCH_DTYPE[dt]
        C, H = ghn.hid, ghn.heads
        ms0, ms1, S, _ = ghn.max_shape
        ncls = ghn.num_classes
        st = ghn._static_device(bp, device)
        self.st = st
        N = bp.total_nodes
        # `shared` (programs of the one-prediction-at-a-time mode): the per-call temporaries are views of grow-only
        # buffers owned by the GHN and shared by the programs of ALL batch plans -- they run one after the other on one
        # stream. A model zoo then needs the workspace of its largest model once instead of one set per architecture
        # (hundreds of MB each at XL, kept alive with the plan), and a cold prediction skips ~20 cudaMallocs (~3 ms).
        # Programs of the overlapped mode run side by side and keep private buffers.
        self.shared = bool(shared) and not train
        ws = _SharedWorkspace.of(ghn, device) if self.shared else None
        E = (lambda *shape, dtype=tdt: ws.empty(shape, dtype)) if ws is not None else \
            (lambda *shape, dtype=tdt: torch.empty(*shape, dtype=dtype, device=device))
        self.keep = []
        self.ops = []            # (stage label, op code, ctypes args struct)
        # ---- node features ----
        if train:                # residual stream of every layer is kept for the backward pass
            self.xs = E(ghn.layers + 1, N, C, dtype=torch.float32)
            self.x = self.xs[0]
        else:
            self.x = E(N, C, dtype=torch.float32)
        t = w['tables']
        self.nf = L.NodeFeaturesArgs(total_nodes=N, hid=C, shape_idx=L.ptr(st['shape_idx']),
                                     embed_op=L.ptr(t['embed_op']), embed_ch=L.ptr(t['embed_ch']),
                                     embed_sp=L.ptr(t['embed_sp']), cent_in=L.ptr(t['cent_in']),
                                     cent_out=L.ptr(t['cent_out']), dist_embed=L.ptr(t['dist_embed']), x=L.ptr(self.x))
        self.ops.append(('node_features', 'node_features', self.nf))
        # ---- Graphormer stack + final LN scattered into the decoder input rows ----
        n_dec = bp.n_conv + bp.n_1d
        self.dec_in = E(max(n_dec, 1), C)
        self.emb = torch.empty(N, C, dtype=torch.float32, device=device) if want_emb else None   # returned to the caller
        self.h, self.qkv, self.ff = E(N, C), E(N, 3 * C), E(N, 4 * C)
        self.h2 = E(N, C)
        self.ln_counters = torch.zeros((N + 127) // 128 + 1, dtype=torch.int32, device=device)
        self.ghn = ghn
        self.ga = L.GraphormerArgs(hid=C, heads=H, layers=ghn.layers, dtype=dt, layers_host=w['layers'],
                                   ln_w=L.ptr(w['ln_w']), ln_b=L.ptr(w['ln_b']), total_nodes=N, x=L.ptr(self.x),
                                   h=L.ptr(self.h), qkv=L.ptr(self.qkv), ff=L.ptr(self.ff), dec_in=L.ptr(self.dec_in),
                                   dec_dtype=act, dst_row=L.ptr(st['dst_row']), emb_f32=L.ptr(self.emb),
                                   ln_counters=L.ptr(self.ln_counters) if FUSE_LN else None, h2=L.ptr(self.h2),
                                   tf32_x3=int(x3))
        if train:
            Lh = ghn.layers
            self.sv = dict(xm=E(Lh, N, C, dtype=torch.float32), h1=E(Lh, N, C), qkv=E(Lh, N, 3 * C), ao=E(Lh, N, C),
                           h2=E(Lh, N, C), u=E(Lh, N, 4 * C), g=E(Lh, N, 4 * C))
            # bf16: the tensor-core attention keeps its softmax statistics for the tensor-core backward
            self.lse2 = E(Lh, H, N, dtype=torch.float32) if act == ops.BF16 else None
            self.ta = L.GraphormerTrainArgs(fwd=self.ga, xs=L.ptr(self.xs), lse2=L.ptr(self.lse2),
                                            **{k_: L.ptr(v_) for k_, v_ in self.sv.items()})
            self.ga = self.ta.fwd            # the struct was copied by value: patch THIS copy in bind_pack
            self.ops.append(('graphormer', 'graphormer_train_fwd', self.ta))
        else:
            # the final LayerNorm (-> decoder input rows) is its own op so that the overlapped mode can order it after
            # the previous call's decoders
            self.ga.skip_final_ln = 1
            D = C // H
            self.fused = None
            if (getattr(ghn, 'fused_graphormer', FUSED_DEFAULT) and dt == ops.BF16 and not x3
                    and w['wstack'] is not None and N <= FUSED_MAX_NODES and C % 64 == 0 and C <= 384
                    and D in (8, 16, 24) and len(bp.plans) <= 256):
                self.fused = ops.FusedGraphormer(C, H, ghn.layers, w['wstack'], w['layers_dev'], N, device, x=self.x)
                self.ops.append(('graphormer', 'graphormer_fused', self.fused.args))
            else:
                self.ops.append(('graphormer', 'graphormer_stack', self.ga))
            self.final_ln = L.LayerNormArgs(rows=N, hid=C, x=L.ptr(self.x), gamma=L.ptr(w['ln_w']), beta=L.ptr(w['ln_b']),
                                            out=L.ptr(self.dec_in), out_dtype=act, dst_row=L.ptr(st['dst_row']),
                                            out_f32=L.ptr(self.emb))
            self.ops.append(('graphormer', 'layernorm', self.final_ln))
        bufs = {}
        self.rts = []

        def gemm_args(a, b, bias, act_, out, out_dtype, problems=None, tiles=None, **kw):
            g = L.GemmArgs(a=L.ptr(a), a_rows=a.shape[0], lda=a.stride(0), b=L.ptr(b), b_rows=b.shape[0],
                           ldb=b.stride(0), k=a.shape[1], in_dtype=dt, d=L.ptr(out), out_dtype=out_dtype,
                           bias=L.ptr(bias), act=act_, tf32_x3=int(x3), **kw)
            if problems is not None:
                g.problems, g.tiles, g.n_tiles = L.ptr(problems), L.ptr(tiles), tiles.shape[0]
            else:
                g.single = L.GemmProblem(a_row0=0, b_row0=0, m=a.shape[0], n=b.shape[0], d_off=0,
                                         ldd=out.stride(0), bias_off=0 if bias is not None else -1)
            return g

        # ---- conv decoder: fc (cropped positions) -> conv.0 -> conv.2 (needed columns only) ----
        R = bp.conv_total_rows
        if R > 0:
            self.h0, self.h1 = E(R, 4 * C), E(R, 8 * C)
            self.wout = E(bp.wout_elems, dtype=torch.float32)
            fc_swap = bp.fc_swap and not x3
            self.ops.append(('dec_fc', 'gemm', gemm_args(self.dec_in, w['fc_w'], w['fc_b'], ops.ACT_RELU, self.h0, act,
                                                         st['fc_problems'],
                                                         st['fc_tiles_swap'] if fc_swap else st['fc_tiles'],
                                                         rowmap=L.ptr(st['fc_rowmap']), swap_ab=int(fc_swap))))
            self.ops.append(('dec_conv0', 'gemm', gemm_args(self.h0, w['c0_w'], w['c0_b'], ops.ACT_RELU, self.h1, act)))
            for g_, _, _ in bp.c2_launches:
                self.ops.append(('dec_conv2', 'gemm',
                                 gemm_args(self.h1, w['c2_w'], w['c2_b'], ops.ACT_NONE, self.wout, ops.F32,
                                           st['c2_%d_problems' % g_], st['c2_%d_tiles' % g_], b_group=g_,
                                           b_group_stride=ms1 if g_ else 0,
                                           block_n=128 if g_ else bp.c2_block_n)))
            bufs[SRC_WOUT] = self.wout
            if bp.clsw_elems:
                self.clsw = E(bp.clsw_elems, dtype=torch.float32)
                for (woff, ld, ii, cnt, coff) in bp.cls_heads:
                    # out[cls][(node, b)] = b_cls[cls] + sum_a W_cls[cls][a] * relu(wout[node][a*i'+b])  (nn.py:757-758):
                    # relu + transpose of the (ms x i') block into a K-major operand, then a tensor-core GEMM
                    rt = E(cnt * ii, ms0)
                    self.keep.append(rt)
                    self.rts.append(rt)
                    ta = L.ReluTransposeArgs(src=self.wout.data_ptr() + woff * 4, ld=ii, src_bs=ld, dst=L.ptr(rt),
                                             dst_dtype=act, rows=ms0, cols=ii, batch=cnt)
                    self.ops.append(('heads_1d', 'relu_transpose', ta))
                    out_view = self.clsw[coff:coff + ncls * cnt * ii].view(ncls, cnt * ii)
                    self.ops.append(('heads_1d', 'gemm', gemm_args(w['cls_w'], rt, w['cls_b'], ops.ACT_NONE, out_view,
                                                                   ops.F32, bias_rows=1, b_dynamic=1)))
                bufs[SRC_CLSW] = self.clsw
        # ---- 1-D decoder (+ classification bias head) ----
        if bp.n_1d > 0:
            d_in = self.dec_in[bp.n_conv:bp.n_conv + bp.n_1d]
            mc = bp.max_ch
            self.hid1, self.d1 = E(bp.n_1d, 2 * C), E(bp.n_1d, 2 * mc, dtype=torch.float32)
            self.ops.append(('heads_1d', 'gemm', gemm_args(d_in, w['d1_w0'], w['d1_b0'], ops.ACT_RELU, self.hid1, act)))
            self.ops.append(('heads_1d', 'gemm', gemm_args(self.hid1, w['d1_w1'], w['d1_b1'], ops.ACT_NONE, self.d1,
                                                           ops.F32)))
            self._d1_ops = [self.ops[-2][2], self.ops[-1][2]]      # depend on the decoder input rows only
            bufs[SRC_D1] = self.d1
            if bp.n_clsb:
                self.clsb = E(2 * bp.n_clsb, ncls, dtype=torch.float32)
                sa = L.GemmSimtArgs(a=self.d1.data_ptr() + (bp.n_1d - bp.n_clsb) * 2 * mc * 4, sam=mc, sak=1,
                                    b=L.ptr(w['bc_w']), sbn=mc, sbk=1, bias=L.ptr(w['bc_b']), d=L.ptr(self.clsb),
                                    sdm=ncls, sdn=1, m=2 * bp.n_clsb, n=ncls, k=mc, relu_a=1, act=ops.ACT_NONE,
                                    batch=1)
                self.ops.append(('heads_1d', 'gemm_simt', sa))
                self._d1_ops.append(sa)
                bufs[SRC_CLSB] = self.clsb
        self.tok = E(max(bp.n_tok_elems, 1), dtype=torch.float32)
        bufs[SRC_TOK] = self.tok
        self.bufs = bufs
        # ---- tile / normalise / scatter into the target parameters ----
        n = len(bp.desc_static)
        self.n_desc = n
        self.desc_host = bp.desc_static.copy()
        base = np.zeros(8, dtype=np.uint64)
        for k_, v in bufs.items():
            base[k_] = v.data_ptr()
        if n:
            self.desc_host['src'] = base[bp.desc_src_buf] + bp.desc_src_off * np.uint64(4)
        self.desc_dev = torch.empty(max(n, 1) * self.desc_host.dtype.itemsize, dtype=torch.uint8, device=device)
        self.last_ptrs = None
        if n:
            self.pred_sumsq = torch.zeros(len(bp.plans), dtype=torch.float64, device=device)
            self.sc = L.ScatterArgs(descs=L.ptr(self.desc_dev), n_descs=n, n_chunks=bp.n_chunks,
                                    chunk_desc=L.ptr(st['chunk_desc']), norm_out=L.ptr(self.pred_sumsq),
                                    n_norm_slots=len(bp.plans))
            self.ops.append(('scatter', 'scatter', self.sc))
        # flat op table for ghn3_run_sequence. `self.ops` stays in dependency order on one lane (the profiled path runs
        # it op by op); the table may reorder the decoder ops onto two lanes:
        #   main: fc -> conv.0 -> large conv.2 launches ........ JOIN -> class heads
        #   aux:  FORK -> 1-D decoder      FORK(after conv.0) -> conv.2 column classes with few tiles
        # The small launches are latency-bound (a K = 8C loop on a handful of tiles): next to the HBM-bound launches
        # they are free, in front of them they cost ~0.1 ms per prediction.
        names = [name for _, name, _ in self.ops]
        i_ln0 = names.index('layernorm') if 'layernorm' in names else None
        stage_of = [st_ for st_, _, _ in self.ops]
        order = [(i, 0) for i in range(len(self.ops))]
        lanes = (not train and i_ln0 is not None and R > 0 and bool(getattr(ghn, 'decoder_lanes', LANES_DEFAULT)))
        if lanes:
            n_sm = torch.cuda.get_device_properties(device).multi_processor_count
            pre = [i for i in range(len(self.ops)) if i <= i_ln0]
            i_fc = [i for i, st_ in enumerate(stage_of) if st_ in ('dec_fc', 'dec_conv0')]
            c2 = [i for i, st_ in enumerate(stage_of) if st_ == 'dec_conv2']
            c2_small = [i for i, (g_, _, tl) in zip(c2, bp.c2_launches) if tl.shape[0] < 2 * n_sm]
            c2_big = [i for i in c2 if i not in c2_small]
            d1 = [i for i in range(len(self.ops)) if self.ops[i][2] in getattr(self, '_d1_ops', [])]
            cls = [i for i, st_ in enumerate(stage_of) if st_ == 'heads_1d' and i not in d1]
            tail = [i for i, st_ in enumerate(stage_of) if st_ == 'scatter']
            if c2_big and (d1 or c2_small):
                order = [(i, 0) for i in pre]
                if d1:
                    order += [('fork', 0)] + [(i, 1) for i in d1]
                order += [(i, 0) for i in i_fc]
                if c2_small:
                    order += [('fork', 0)] + [(i, 1) for i in c2_small]
                order += [(i, 0) for i in c2_big] + [('join', 0)] + [(i, 0) for i in cls] + [(i, 0) for i in tail]
                assert sorted(i for i, _ in order if not isinstance(i, str)) == list(range(len(self.ops)))
        self.seq = (L.SeqOp * len(order))()
        self.seq_len = len(order)
        for k, (i, lane) in enumerate(order):
            if isinstance(i, str):
                self.seq[k].op = 24 if i == 'fork' else 25
                self.seq[k].args = None
            else:
                self.seq[k].op = self.OP[self.ops[i][1]]
                self.seq[k].args = ct.cast(ct.pointer(self.ops[i][2]), ct.c_void_p)
            self.seq[k].lane = lane
        pos = {i: k for k, (i, _) in enumerate(order) if not isinstance(i, str)}
        self.seq_ln = pos[i_ln0] if i_ln0 is not None else None          # table position of the final LayerNorm
        self.bound_pack = None
        self.bwd = None
        # CUDA-graph replay (inference programs): the graph pack's device arrays are mirrored into buffers this
        # program owns, so that every kernel argument is fixed and the captured sequences stay valid for any pack
        self.use_graphs = (not train and getattr(self, 'fused', None) is None
                           and bool(getattr(ghn, 'cuda_graphs', GRAPHS_DEFAULT)))
        self.mirror = None
        self.graphs = {}
        self.graph_key = None
        self.runs = 0
        self._finalizer = weakref.finalize(self, _destroy_sequences, self.graphs)
        self._finalizer.atexit = False            # the CUDA context may already be gone at interpreter exit

    def bind_pack(self, pack):
        if pack is self.bound_pack:
            return
        assert pack.total_nodes == self.bp.total_nodes
        if getattr(pack, 'op_dev', None) is None:
            raise RuntimeError('internal: graph pack has no op ids')
        nf, ga = self.nf, self.ga
        lut = self.ghn._lut(self.w, pack.cutoff)
        ga.n_graphs, ga.max_nodes, ga.lut_size = pack.n_graphs, pack.max_nodes, lut.shape[1]
        ga.lut = L.ptr(lut)
        if self.use_graphs:
            blob = pack._blob.dev
            m = self.mirror
            if m is None or m['blob'].numel() != blob.numel() or m['layout'] != pack.derived_layout:
                self.drop_graphs()
                m = self.mirror = {'blob': torch.empty_like(blob), 'derived': torch.empty_like(pack.derived),
                                   'layout': pack.derived_layout}
                off = {name: o for name, o, _ in pack._blob.parts}
                b0 = m['blob'].data_ptr()
                deg_in, deg_out, dist0, pair = ops.derived_views(m['derived'], pack.derived_layout)
                nf.op, nf.deg_in, nf.deg_out, nf.dist0 = b0 + off['op'], L.ptr(deg_in), L.ptr(deg_out), L.ptr(dist0)
                ga.node_off, ga.mat_off, ga.pair = b0 + off['node_off'], b0 + off['mat_off'], L.ptr(pair)
                m['copies'] = (L.MemcpyArgs(dst=L.ptr(m['blob']), bytes=blob.numel()),
                               L.MemcpyArgs(dst=L.ptr(m['derived']), bytes=pack.derived.numel()))
                m['seq'] = (L.SeqOp * 2)()
                for i, a in enumerate(m['copies']):
                    m['seq'][i].op = 23
                    m['seq'][i].args = ct.cast(ct.pointer(a), ct.c_void_p)
            m['copies'][0].src, m['copies'][1].src = L.ptr(blob), L.ptr(pack.derived)
            m['pending'] = True                   # copied by run() on the stream the kernels are launched on
            self.graph_key = (ga.lut, pack.n_graphs, pack.max_nodes)
        else:
            nf.op, nf.deg_in, nf.deg_out, nf.dist0 = L.ptr(pack.op_dev), L.ptr(pack.deg_in), L.ptr(pack.deg_out), \
                L.ptr(pack.dist0)
            ga.node_off, ga.mat_off = L.ptr(pack.d['node_off']), L.ptr(pack.d['mat_off'])
            ga.pair = L.ptr(pack.pair)
        if getattr(self, 'fused', None) is not None:
            self.fused.bind(pack, lut)
        self.bound_pack = pack

    def drop_graphs(self):
        _destroy_sequences(self.graphs)

    def launch(self, i0, i1, stream_ptr, what, high_priority=False):
        """Enqueues ops[i0:i1) on the stream: as a captured CUDA graph from the program's second run on (the first run
        issues the kernels directly -- it also sets their function attributes), kernel by kernel otherwise."""
        lib = L.load()
        if i1 <= i0:
            return
        m = self.mirror
        if m is not None and m.get('pending'):    # mirror this call's graph pack first (stream order = data order)
            L.check(lib.ghn3_run_sequence(m['seq'], 2, ct.c_void_p(stream_ptr)), 'ghn3_run_sequence (pack mirror)')
            m['pending'] = False
        at = ct.c_void_p(ct.addressof(self.seq) + i0 * ct.sizeof(L.SeqOp))
        if self.use_graphs and self.runs > 0:
            key = (i0, i1, bool(high_priority), L._pdl_state[0], L._cap_state[0]) + tuple(self.graph_key or ())
            g = self.graphs.get(key)
            if g is None:
                h = ct.c_void_p()
                rc = lib.ghn3_sequence_capture(at, i1 - i0, int(bool(high_priority)), ct.byref(h))
                if rc != 0:                       # capture not possible here: keep launching kernel by kernel
                    self.use_graphs = False
                    self.capture_error = lib.ghn3_last_error().decode(errors='replace')
                    if self.ghn.debug_level:
                        log('ghn3_b200: CUDA-graph capture failed (%s); launching kernels one by one' %
                            self.capture_error)
                else:
                    g = self.graphs[key] = h.value
                    while len(self.graphs) > 16:
                        old = next(iter(self.graphs))
                        lib.ghn3_sequence_destroy(ct.c_void_p(self.graphs.pop(old)))
            if g is not None:
                L.check(lib.ghn3_sequence_launch(ct.c_void_p(g), ct.c_void_p(stream_ptr)),
                        'ghn3_sequence_launch (%s)' % what)
                return
        L.check(lib.ghn3_run_sequence(at, i1 - i0, ct.c_void_p(stream_ptr)), 'ghn3_run_sequence (%s)' % what)

    def refresh_targets(self, weight_norm):
        """Re-reads the addresses of the target parameters; uploads the descriptor table only if one moved."""
        n = self.n_desc
        if n == 0:
            return
        device = self.device
        ptrs = np.empty(n, dtype=np.uint64)
        for i, (module, attr, shape, view) in enumerate(self.bp.desc_targets):
            p = module._parameters.get(attr)
            if p is None:
                p = getattr(module, attr)
            if not isinstance(p, torch.Tensor):
                raise RuntimeError('ghn3_b200: target %s.%s holds a shape, not a tensor: parameter-free (light) modules '
                                   'can only receive predictions with keep_grads=True (reference nn.py:530-546)'
                                   % (type(module).__name__, attr))
            ptrs[i] = p.data_ptr()
        if self.last_ptrs is not None and np.array_equal(ptrs, self.last_ptrs):
            return
        for i, (module, attr, shape, view) in enumerate(self.bp.desc_targets):
            p = getattr(module, attr)
            if not isinstance(p, torch.Tensor):
                raise RuntimeError('ghn3_b200: target %s.%s holds a shape, not a tensor: parameter-free (light) modules '
                                   'can only receive predictions with keep_grads=True (reference nn.py:530-546)'
                                   % (type(module).__name__, attr))
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                # reference semantics (nn.py:548): param.data is replaced by a tensor on the GHN's device
                p.data = torch.empty(tuple(p.shape), dtype=torch.float32, device=device)
                ptrs[i] = p.data_ptr()
        if getattr(self, 'ev_scatter', None) is not None:      # an overlapped scatter may still read the table
            torch.cuda.current_stream().wait_event(self.ev_scatter)
        desc = self.desc_host
        desc['dst'] = ptrs + self.bp.desc_dst_shift
        if not weight_norm:
            desc['mode'] = np.where(desc['mode'] == 3, 3, 0)
            desc['scale'] = 1.0
        self.desc_dev.copy_(torch.from_numpy(desc.view(np.uint8).reshape(-1)))
        self.last_ptrs = ptrs

    def point_descriptors_at(self, pred, out_meta, weight_norm):
        """Training path: the scatter writes into slices of the flat buffer `pred` (one slice per parameter)."""
        n = self.n_desc
        if n == 0:
            return
        base = pred.data_ptr()
        offs = np.array([out_meta['slices'][oi][0] for oi in out_meta['desc_out']], dtype=np.uint64) * np.uint64(4)
        desc = self.desc_host
        desc['dst'] = np.uint64(base) + offs + self.bp.desc_dst_shift
        if not weight_norm:
            desc['mode'] = np.where(desc['mode'] == 3, 3, 0)
            desc['scale'] = 1.0
        self.desc_dev.copy_(torch.from_numpy(desc.view(np.uint8).reshape(-1)))
        self.last_ptrs = None

    def run(self, prof=None, overlap=False):
        stream = L.current_stream()
        lib = L.load()
        cur = torch.cuda.current_stream()
        ev_prev = getattr(self, 'ev_scatter', None)
        draw_tok = self.bp.n_tok_elems > 0       # ViT class-token rows: fresh N(0, 0.02) draws as in nn.py:446
        if not (prof is None and overlap and self.n_desc):
            self.ghn.flush_all()                 # earlier overlapped calls (any program slot) write the same targets
        if prof is None and not (overlap and self.n_desc):
            if ev_prev is not None:              # an earlier overlapped call: its scatter still reads our buffers
                cur.wait_event(ev_prev)
                self.ev_scatter = None
            if draw_tok:
                self.tok.normal_(mean=0.0, std=0.02)
            self.launch(0, self.seq_len, stream, 'prediction')
            self.runs += 1
            return
        if prof is None:
            # Overlapped form: [node features, Graphormer] -> wait for the PREVIOUS call's scatter (it reads the decoder
            # outputs this call is about to overwrite) -> decoders -> scatter on a side stream. Successive predictions
            # are independent, so the HBM-write-bound scatter of call k runs under the latency-bound Graphormer of
            # call k+1. The caller must use GHN3.flush() before touching the predicted parameters.
            if getattr(self, 'side', None) is None:
                # the Graphormer stack runs on a HIGH-priority stream (one per program slot), decoders + scatter on a
                # normal one shared by all programs (their scatters write the same targets: stream order = call order):
                # when the two compete for SM slots the block scheduler serves the latency-bound chain first
                self.hi, self.side, self.sc = self.ghn._overlap_streams(self.device, self.slot)
                self.i_ln = self.seq_ln
                self.ev_a = torch.cuda.Event()
                self.ev_dec = None
            hi, side, sc = self.hi, self.side, self.sc
            n_ops, i_ln, i_sc = self.seq_len, self.i_ln, self.seq_len - 1
            split = self.ghn.__dict__.get('overlap_decoders', True)      # False: only the scatter leaves the main chain
            run = lambda i0, i1, st_, what: self.launch(i0, i1, st_.cuda_stream, what, high_priority=st_ is hi)
            ev_in = torch.cuda.Event()
            ev_in.record(cur)                    # the graph pack was uploaded / derived on the caller's stream
            hi.wait_event(ev_in)
            self.bound_pack.record_stream(hi)
            run(0, i_ln, hi, 'graphormer')
            # the final LayerNorm overwrites the decoder input rows: the previous call's decoders must have read them
            if self.ev_dec is not None:
                hi.wait_event(self.ev_dec)
            run(i_ln, i_ln + 1, hi, 'final layernorm')
            if not split:
                if ev_prev is not None:
                    hi.wait_event(ev_prev)
                run(i_ln + 1, i_sc, hi, 'decoders')
            self.ev_a.record(hi)
            side.wait_event(self.ev_a)
            if split:
                if ev_prev is not None and sc is not side:
                    side.wait_event(ev_prev)     # this program's previous scatter still reads its decoder outputs
                run(i_ln + 1, i_sc, side, 'decoders')
            ev_dec = torch.cuda.Event()
            ev_dec.record(side)
            self.ev_dec = ev_dec
            # scatters of all programs on ONE stream (they write the same targets: stream order = call order)
            if sc is not side:
                sc.wait_event(ev_dec)
            if draw_tok:
                with torch.cuda.stream(sc):
                    self.tok.normal_(mean=0.0, std=0.02)
            run(i_sc, n_ops, sc, 'scatter')
            ev = torch.cuda.Event()
            ev.record(sc)
            self.ev_scatter = ev
            self.runs += 1
            return
        if ev_prev is not None:
            cur.wait_event(ev_prev)
            self.ev_scatter = None
        if draw_tok:
            self.tok.normal_(mean=0.0, std=0.02)

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            prof.append((name, ev))
        if self.mirror is not None and self.mirror.get('pending'):
            L.check(lib.ghn3_run_sequence(self.mirror['seq'], 2, ct.c_void_p(stream)), 'ghn3_run_sequence (pack mirror)')
            self.mirror['pending'] = False
        mark('start')
        per_op = getattr(prof, 'per_op', False)       # per-launch events (bench.py's roofline leg)
        for i, (stage, name, args) in enumerate(self.ops):
            L.call(name, args, stream)
            if per_op:
                mark('%s#%d' % (stage, i))
            elif i + 1 == len(self.ops) or self.ops[i + 1][0] != stage:
                mark(stage)


# ----------------------------------------------------------------------------------------------------------------
# Checkpoints released without a 'config' entry: the GHN hyper-parameters are read off the tensors. One rule per
# hyper-parameter: (name test, value from the tensor); the checkpoint schema is the reference's (nn.py:59-99).
_CONFIG_RULES = (
    ('num_classes', lambda n: 'class_layer_predictor' in n, lambda p: p.shape[0]),
    ('layernorm', lambda n: n.endswith('ln.weight'), lambda p: True),
    ('hid', lambda n: n.endswith('embed.weight'), lambda p: p.shape[-1]),
    ('max_ch', lambda n: n.endswith('decoder.conv.2.weight'), lambda p: int(p.shape[0] ** 0.5)),
    ('spatial', lambda n: n.endswith('shape_enc.embed_spatial.weight'), lambda p: 11 if p.shape[0] == 9 else 16),
    ('pretrained', lambda n: 'centrality_embed_in' in n and 'gnn.' not in n, lambda p: True),
)


def infer_config(state_dict, overrides=None):
    """GHN3 constructor arguments from a bare state_dict. `overrides` (popped): defaults for what no tensor settles."""
    kw = overrides if overrides is not None else {}
    found = {'num_classes': kw.pop('num_classes', 10), 'hid': kw.pop('hid', 32), 'layernorm': kw.pop('layernorm', False),
             'pretrained': kw.pop('pretrained', False), 'max_ch': kw.pop('max_shape', 64), 'spatial': None}
    layers = kw.pop('layers', 0)
    for name, p in state_dict.items():
        for key, test, value in _CONFIG_RULES:
            if test(name):
                found[key] = value(p)
        layers += int(name.endswith('ln1.weight') and 'gnn.' in name)      # one ln1 per Graphormer layer
    s = found['spatial'] or (16 if found['num_classes'] >= 1000 else 11)   # decoder grid: 16 for ImageNet GHNs
    ms = found['max_ch']
    return {'hid': found['hid'], 'max_shape': ms if isinstance(ms, tuple) else (ms, ms, s, s),
            'num_classes': found['num_classes'], 'heads': 16 if found['hid'] > 64 else 8, 'layers': layers,
            'weight_norm': True, 've': True, 'layernorm': found['layernorm'], 'pretrained': found['pretrained']}


def from_pretrained(ghn3_name='ghn3xlm16.pt', **kwargs):
    """
    Loads a GHN-3 checkpoint (reference nn.py:31-125): a local file holding {'state_dict', 'config'?} or a bare
    state_dict; the GHN config is inferred from tensor names/shapes when absent. Hugging Face download is attempted
    only if the file does not exist locally and huggingface_hub is importable.
    """
    import os
    assert ghn3_name is not None, 'GHN ckpt must be specified.'
    state_dict, ghn_config = None, None
    if not os.path.exists(ghn3_name):
        try:
            import joblib
            from huggingface_hub import hf_hub_download
            state_dict = joblib.load(hf_hub_download(repo_id='SamsungSAILMontreal/ghn3', filename=ghn3_name))
        except Exception as e:
            raise FileNotFoundError('cannot load GHN checkpoint %s: not a local file and the hub download failed (%s)'
                                    % (ghn3_name, e))
    else:
        try:
            state_dict = torch.load(ghn3_name, map_location='cpu', weights_only=False)
        except Exception:
            import joblib                    # the released checkpoints are joblib pickles of a state_dict (nn.py:49)
            state_dict = joblib.load(ghn3_name)
        if 'config' in state_dict:
            ghn_config = state_dict['config']
        if 'state_dict' in state_dict:
            state_dict = state_dict['state_dict']
    if any(k.find('gnn.gru.') >= 0 for k in state_dict):
        raise NotImplementedError('GHN-2 checkpoints are out of scope of ghn3_b200')
    compute_dtype = kwargs.pop('compute_dtype', 'bf16')
    if ghn_config is None:
        ghn_config = infer_config(state_dict, kwargs)
    else:
        ghn_config = dict(ghn_config)
        ghn_config.pop('is_ghn2', None)
    cache_compute_copy = kwargs.pop('cache_compute_copy', False)
    ghn = GHN3(**ghn_config, compute_dtype=compute_dtype, **kwargs)
    ghn.load_state_dict(state_dict)
    ghn.fix_embed_layers()
    if cache_compute_copy and os.path.exists(ghn3_name):
        # the compute-dtype copy of the GEMM weights (bf16: 1.3 GB instead of 2.6 GB at ghn3xlm16, decoder
        # fc already repacked position-major) is written beside the checkpoint on first use and read back afterwards
        ghn._compute_cache_path = '%s.%s.cache' % (ghn3_name, compute_dtype)
        ghn._compute_cache_key = (os.path.getsize(ghn3_name), int(os.path.getmtime(ghn3_name)))
    return ghn


def param_norm(model, out=None):
    """Total L2 norm of a model's parameters, computed on the device by ghn3_sumsq (the norm_check metric of
    nn.py:783-797). Returns a 1-element float64 CUDA tensor (no host sync). The (pointer, numel) table of the model
    is cached on the model and re-uploaded only when a parameter moved."""
    cache = model.__dict__.get('_ghn3_b200_norm')
    if cache is None:
        cache = {'params': [p for p in model.parameters()], 'ptrs': None, 'meta': None}
        model.__dict__['_ghn3_b200_norm'] = cache
    ps = cache['params']
    ptrs = [p.data_ptr() for p in ps]
    if ptrs != cache['ptrs']:
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError('ghn3_b200.param_norm: parameters must be contiguous fp32 CUDA tensors')
        meta = np.array([ptrs, [p.numel() for p in ps]], dtype=np.int64)
        cache['meta'] = torch.from_numpy(meta).to(ps[0].device)
        cache['ptrs'] = ptrs
        cache['args'] = L.SumsqArgs(ptrs=cache['meta'][0].data_ptr(), numels=cache['meta'][1].data_ptr(), n=len(ps))
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=ps[0].device)
    a = cache['args']
    a.out = out.data_ptr()
    L.call('sumsq', a, L.current_stream())
    return out.sqrt_()
