"""
Host-side planning for the CUDA path: everything about a (graph, target model, GHN config) triple that does not
depend on the GHN weights or on tensor addresses is computed here once, in numpy, and cached:

  * the node -> target-parameter mapping and the shape-group keys      (reference ghn3/nn.py:594-692)
  * the ShapeEncoder look-up indices of every node                      (ppuda ShapeEncoder)
  * the decoder work list: which rows of the decoder weights each node needs (ghn3/nn.py:735-762), as grouped-GEMM
    problems / tiles for ghn3_gemm
  * one scatter descriptor per predicted tensor                          (ghn3/nn.py:422-506, 554-592, 508-552)

Nothing in this file touches the GPU or does floating-point work on parameters.
"""
from collections import OrderedDict

import math

import numpy as np
import torch
import torch.nn as nn

from ._lib import fastdiv
from .weights import channel_bins, spatial_bins

try:
    from torchvision.models.vision_transformer import Encoder as _VitEncoder
except Exception:  # pragma: no cover
    _VitEncoder = ()

_ATTR_NAMES = frozenset(('weight', 'bias', 'in_proj_weight', 'in_proj_bias', 'pos_embedding'))
_PARAM_ATTRS = (('weight', '.weight', True), ('bias', '.bias', False), ('in_proj_weight', '.in_proj_weight', True),
                ('in_proj_bias', '.in_proj_bias', False), ('pos_embedding', '.pos_embedding.weight', True))


# ----------------------------------------------------------------------------------------------------------------
# target-model introspection (ppuda named_layered_modules / get_cell_ind)
# ----------------------------------------------------------------------------------------------------------------
def cell_index(name, n_cells):
    pos = name.find('cells.')
    if pos >= 0:
        return int(name[pos + 6:].split('.', 1)[0])
    if name.startswith(('classifier', 'auxiliary')):
        return n_cells - 1
    if n_cells == 1 or name.startswith(('stem', 'pos_enc')):
        return 0
    return None


_cls_attr_memo = {}


def _class_level_attrs(cls):
    """Names of _ATTR_NAMES defined on the class (or a base) rather than per instance: properties, class defaults."""
    hit = _cls_attr_memo.get(cls)
    if hit is None:
        hit = _cls_attr_memo[cls] = frozenset(a for a in _ATTR_NAMES if hasattr(cls, a))
    return hit


def layered_modules(model):
    """[{param_name: entry}] per cell; entry = dict(param_name, module, is_w, sz)."""
    if hasattr(model, 'module'):
        model = model.module
    n_cells = getattr(model, '_n_cells', 1)
    cells = [OrderedDict() for _ in range(n_cells)]
    for name, mod in model.named_modules():
        hits = []
        md = mod.__dict__
        pd = md.get('_parameters', {})
        cls_attrs = _class_level_attrs(type(mod))
        if not pd and not cls_attrs and _ATTR_NAMES.isdisjoint(md):
            continue                       # containers, activations, ...: nothing to predict
        for attr, suffix, is_w in _PARAM_ATTRS:
            # registered parameter, then plain attribute (the shape lists of parameter-free modules, tensors set by a
            # keep_grads prediction); nn.Module.__getattr__ for a missing name costs ~2 us x 8 names x every module
            p = pd.get(attr)
            if p is None:
                p = md.get(attr)
                if p is None and attr in cls_attrs:
                    p = getattr(mod, attr, None)
            if p is None or isinstance(p, bool) or not isinstance(p, (torch.Tensor, list, tuple)):
                continue
            hits.append((name + suffix, p, is_w))
        if not hits:
            continue
        ci = cell_index(name, n_cells)
        ci = 0 if ci is None else ci
        for key, p, is_w in hits:
            cells[ci][key] = {'param_name': key, 'module': mod, 'is_w': is_w,
                              'sz': tuple(p) if isinstance(p, (list, tuple)) else tuple(p.shape)}
    return cells


def param_attr(module, is_w):
    """Attribute that receives the predicted tensor (nn.py:519-524)."""
    if isinstance(module, nn.MultiheadAttention):
        return 'in_proj_weight' if is_w else 'in_proj_bias'
    if _VitEncoder and isinstance(module, _VitEncoder):
        return 'pos_embedding'
    return 'weight' if is_w else 'bias'


# ----------------------------------------------------------------------------------------------------------------
# ShapeEncoder index tables
# ----------------------------------------------------------------------------------------------------------------
class ShapeIndexer:
    """Vectorised version of the ShapeEncoder lookup dictionaries: dense tables indexed by the size value."""

    _cache = {}

    def __init__(self, num_classes, max_shape):
        ch, sp = channel_bins(num_classes), spatial_bins(max_shape)
        self.n_ch, self.n_sp = len(ch), len(sp)
        self.ch_table = self._table(ch, {c: 8 for c in range(4, 8)})
        self.sp_table = self._table(sp, {2: 3})
        self.dummy = np.array([self.n_ch, self.n_ch, self.n_sp, self.n_sp], dtype=np.int32)
        self._memo = {}

    @staticmethod
    def _table(bins, alias):
        top = int(bins[-1])
        tab = np.empty(top + 1, dtype=np.int32)
        pos = {int(b): i for i, b in enumerate(bins)}
        for v in range(top + 1):
            if v in pos:
                tab[v] = pos[v]
            elif v in alias:
                tab[v] = pos[alias[v]]
            else:
                tab[v] = int(np.argmin(np.abs(bins - v)))     # nearest bin, first on ties
        tab[0] = pos[top]                                    # 0 is not in the dictionaries -> "largest" rule
        return tab

    @classmethod
    def get(cls, num_classes, max_shape):
        key = (num_classes, tuple(max_shape))
        if key not in cls._cache:
            cls._cache[key] = cls(num_classes, max_shape)
        return cls._cache[key]

    def lookup(self, sz):
        """int32[4] index row of a tensor shape (memoised: a model zoo repeats a few hundred shapes)."""
        sz = tuple(int(v) for v in sz)
        hit = self._memo.get(sz)
        if hit is None:
            hit = self._memo[sz] = self._lookup(sz)
        return hit

    def _lookup(self, sz):
        if len(sz) == 1:
            sz = (sz[0], 1)
        if len(sz) == 2:
            sz = (sz[0], sz[1], 1, 1)
        if len(sz) == 3:
            sz = (sz[0], sz[1], sz[2], 1)
        out = np.empty(4, dtype=np.int32)
        for i in range(4):
            tab = self.ch_table if i < 2 else self.sp_table
            v = sz[i]
            out[i] = tab[v] if 0 < v < len(tab) else tab[len(tab) - 1]
        return out


# ----------------------------------------------------------------------------------------------------------------
# per-model plan
# ----------------------------------------------------------------------------------------------------------------
KIND_CONV, KIND_CLS_W, KIND_1D, KIND_CLS_B, KIND_3D = 0, 1, 2, 3, 4


class NodeTask:
    """One decoded node: where its prediction comes from and which target tensors it fills."""
    __slots__ = ('node', 'kind', 'key', 'entry', 'o_need', 'i_need', 'win', 'interp')


class ModelPlan:
    """Static plan of one (graph, model) pair under one GHN config."""

    def __init__(self, graph, model, cfg, predict_class_layers=True):
        self.n_nodes = graph.n_nodes
        self.cfg = cfg
        ms = cfg['max_shape']
        S = ms[2]
        cells = layered_modules(model)
        indexer = ShapeIndexer.get(cfg['num_classes'], ms)
        self.shape_idx = np.tile(indexer.dummy, (self.n_nodes, 1))
        self.groups = OrderedDict()                 # key -> [node ids]  (first-appearance order, nn.py:677-680)
        self.tasks = []
        self.n_tensors = 0
        self.n_params = 0
        self._cells = cells
        self._matched = set()
        for cell_id, cell in enumerate(graph.node_info):
            for (node_ind, p_, name, sz, last_weight, last_bias) in cell:
                p_name = p_ if p_.endswith(('.weight', '.bias', 'in_proj_weight', 'in_proj_bias')) else p_ + '.weight'
                entry = cells[cell_id].get(p_name)
                if entry is None:
                    entry = cells[cell_id].get(p_name.replace('to_qkv', 'attn.to_qkv').replace('to_out', 'attn.to_out'))
                if entry is None:
                    if sz is not None:
                        self.shape_idx[node_ind] = indexer.lookup(sz)
                    continue
                tsz = entry['sz']
                self._matched.add((cell_id, entry['param_name']))
                self.shape_idx[node_ind] = indexer.lookup(tsz)
                key = self._group_key(tsz, ms, last_weight, last_bias)
                self.groups.setdefault(key, []).append(node_ind)
                t = NodeTask()
                t.node, t.key, t.entry = node_ind, key, entry
                if len(key) == 2 and key[1] > 0:
                    t.kind, t.o_need, t.i_need = KIND_CLS_W, ms[0], key[1]
                    self._window(t, 1, 1, S)
                elif len(key) == 2:
                    t.kind = KIND_CLS_B if key[1] < 0 else KIND_1D
                elif len(key) == 3:
                    # 3-D tensors that are not positional encodings (nn.py:287-289,672: "e.g. layer_scale"): the
                    # 2*ms values of decoder_1d, first dimension cropped, last dimension tiled. The reference's
                    # _tile_params only yields the target shape for (o <= 2 ms, 1, k) (its repeat() calls are 4-D).
                    if tsz[1] != 1 or tsz[0] > 2 * max(ms[0], ms[1]):
                        raise NotImplementedError('3-D target of shape %s: the reference (nn.py:438-488) only handles '
                                                  '(o <= 2*max_shape, 1, k)' % (tuple(tsz),))
                    t.kind = KIND_3D
                else:
                    t.kind, t.o_need, t.i_need = KIND_CONV, min(key[0], ms[0]), min(key[1], ms[1])
                    self._window(t, key[2], key[3], S)
                is_cls = t.kind in (KIND_CLS_W, KIND_CLS_B)
                if is_cls and not predict_class_layers:
                    continue
                self.tasks.append(t)
                n_t = 2 if (len(tsz) == 1 and entry['is_w'] and getattr(entry['module'], 'bias', None) is not None) \
                    else 1
                self.n_tensors += n_t
                self.n_params += math.prod(int(v) for v in tsz) * n_t

    def prune_unmatched(self):
        """reduce_graph=True (nn.py:684-690): parameters of modules that no graph node refers to are set to None."""
        for cell_id, cell in enumerate(self._cells):
            for name, e in cell.items():
                if (cell_id, name) in self._matched or not e['is_w']:
                    continue
                mod = e['module']
                mod.weight = None
                if hasattr(mod, 'bias') and mod.bias is not None:
                    mod.bias = None

    @staticmethod
    def _group_key(sz, ms, last_weight, last_bias):
        def min_sz(j):                                   # nn.py:652-660
            n = min(sz[j], ms[j])
            if n % 3 == 0:
                n = n // 3 * 4
            if n >= ms[j] / 2:
                n = ms[j]
            return n
        if len(sz) == 1:
            return (min_sz(0), -1) if last_bias else (min_sz(0), 0)
        if last_weight:
            return (min_sz(0), min_sz(1))
        if len(sz) == 2:
            return (min_sz(0), min_sz(1), 1, 1)
        if len(sz) == 3:
            if sz[0] == 1 and min(sz[1:]) > 1:
                s = int(np.floor(sz[1] ** 0.5))
                return (1, sz[2], s, s)
            return (min_sz(0), min_sz(1), min_sz(2))
        return (min_sz(0), min_sz(1), sz[2], sz[3])

    @staticmethod
    def _window(t, kh, kw, S):
        """Centred crop of the S x S decoder grid (nn.py:742-747) -> (y0, y1, x0, x1); bilinear flag (nn.py:751)."""
        off = S // 2
        y0, y1 = max(0, off - kh // 2), min(S, off + (kh + 1) // 2)
        x0, x1 = max(0, off - kw // 2), min(S, off + (kw + 1) // 2)
        t.win = (y0, y1, x0, x1)
        t.interp = min(kh, kw) > min(min(S, kh), min(S, kw))


_scale_memo = {}


def scale_for(sz):
    """Fan-in normalisation factor of nn.py:562-583 for a tensor of shape sz (>1-D); rounded once to fp32."""
    key = tuple(int(v) for v in sz)
    hit = _scale_memo.get(key)
    if hit is None:
        hit = _scale_memo[key] = _scale_for(key)
    return hit


def _scale_for(sz):
    if len(sz) > 2 and sz[2] >= 11 and sz[0] == 1:
        return 1.0
    no_relu = len(sz) > 2 and (sz[1] == 1 or sz[2] < sz[3])
    beta = 1.0 if no_relu else 2.0
    return float(np.float32((beta / float(math.prod(int(v) for v in sz[1:]))) ** 0.5))


# ----------------------------------------------------------------------------------------------------------------
# batch plan: merges the per-model plans into decoder GEMM problems and scatter descriptors
# ----------------------------------------------------------------------------------------------------------------
PROBLEM_DT = np.dtype([('a_row0', 'i4'), ('b_row0', 'i4'), ('m', 'i4'), ('n', 'i4'), ('d_off', 'i8'), ('ldd', 'i4'),
                       ('bias_off', 'i4')])
DESC_DT = np.dtype([('dst', 'u8'), ('src', 'u8'), ('numel', 'i8'), ('chunk0', 'i8'), ('t1', 'i4'), ('t2', 'i4'),
                    ('t3', 'i4'), ('so', 'i4'), ('si', 'i4'), ('ld', 'i4'), ('ca', 'i4'), ('ra', 'i4'),
                    ('kh_src', 'i4'), ('kw_src', 'i4'), ('cy', 'i4'), ('cx', 'i4'), ('scale', 'f4'), ('mode', 'i4')] +
                   [(n_, 'u4') for n_ in ('m_t1', 's_t1', 'm_t2', 's_t2', 'm_t3', 's_t3', 'm_so', 's_so', 'm_si', 's_si')] +
                   [('norm_slot', 'i4'), ('reserved', 'i4')])
assert PROBLEM_DT.itemsize == 32 and DESC_DT.itemsize == 136
SCATTER_CHUNK = 16384
SRC_WOUT, SRC_D1, SRC_CLSW, SRC_CLSB, SRC_TOK = 0, 1, 2, 3, 4      # which device buffer a descriptor reads
import os as _os
C2_DENSE_BLOCK_N = int(_os.environ.get('GHN3_C2_BLOCK_N', '128'))    # N tile of the dense conv.2 launch


def tiles_for(problems, block_m=128, block_n=128):
    """(problem, M tile, N tile, 0) rows for every tile of every problem, N tiles of one M tile adjacent."""
    if len(problems) == 0:
        return np.zeros((0, 4), dtype=np.int32)
    mt = -(-problems['m'].astype(np.int64) // block_m)
    nt = -(-problems['n'].astype(np.int64) // block_n)
    cnt = mt * nt
    total = int(cnt.sum())
    p = np.repeat(np.arange(len(problems), dtype=np.int64), cnt)
    local = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    ntp = nt[p]
    t = np.zeros((total, 4), dtype=np.int32)
    t[:, 0] = p
    t[:, 1] = local // np.maximum(ntp, 1)
    t[:, 2] = local % np.maximum(ntp, 1)
    return t


def fastdiv_array(d):
    """Vectorised _lib.fastdiv: (mul, shift) arrays with n // d == (n * mul >> 32) >> shift, mul == 0 for d <= 1."""
    d = np.asarray(d, dtype=np.int64)
    big = d > 1
    dd = np.where(big, d, 2)
    s = np.ceil(np.log2(dd.astype(np.float64))).astype(np.int64)
    s += (1 << s) < dd                         # guard against rounding of log2 just below a power of two
    s -= ((1 << (s - 1)) >= dd) & (s > 1)
    mul = -((-(np.int64(1) << (31 + s))) // dd)
    assert bool((mul[big] < (1 << 32)).all())
    return np.where(big, mul, 0).astype(np.uint32), np.where(big, s - 1, 0).astype(np.uint32)


class BatchPlan:
    """Decoder / scatter work of a batch of (graph, model) pairs. Node ids are global (packed) indices."""

    def __init__(self, plans, cfg):
        self.plans = plans
        self.cfg = cfg
        C = cfg['hid']
        ms0, ms1, S, _ = cfg['max_shape']
        ncls = cfg['num_classes']
        max_ch = max(ms0, ms1)
        offs = np.concatenate([[0], np.cumsum([p.n_nodes for p in plans])]).astype(np.int64)
        self.node_off = offs
        self.total_nodes = int(offs[-1])
        self.shape_idx = np.concatenate([p.shape_idx for p in plans]) if plans else np.zeros((0, 4), np.int32)

        conv, one_d = [], []
        for b, p in enumerate(plans):
            for t in p.tasks:
                (conv if t.kind in (KIND_CONV, KIND_CLS_W) else one_d).append((b, t))

        # ---- conv-decoder nodes: sort by column class, then cls flag, then crop window ----
        def conv_order(bt):
            b, t = bt
            return (-(t.o_need * t.i_need), -t.o_need, -t.i_need, t.kind == KIND_CLS_W, t.win, b, t.node)
        conv.sort(key=conv_order)
        self.conv = conv
        self.n_conv = len(conv)
        dst_row = np.full(self.total_nodes, -1, dtype=np.int32)         # row of each node in the decoder inputs
        fc_probs, c2_probs = [], []
        c2_grouped = {}                  # i' -> problems reading the compact o' x i' column set through a 3-D TMA box
        self.conv_rows = []              # (row0 in h0/h1, P, kwp, khp) per conv node
        self.seg_of = []                 # wout element offset of each conv node's first row, ld
        row = 0
        wout_elems = 0
        i = 0
        self.cls_heads = []              # (wout_off, ld, i_need, n_nodes, clsw_out_off)
        self.segments = []               # (o', i', first row, rows, wout offset): one per column class (training path)
        clsw_elems = 0
        while i < len(conv):
            o, ii = conv[i][1].o_need, conv[i][1].i_need
            j = i
            while j < len(conv) and (conv[j][1].o_need, conv[j][1].i_need) == (o, ii):
                j += 1
            ld = o * ii
            seg_row0, seg_base = row, wout_elems
            k = i
            while k < j:                 # sub-groups of equal (cls flag, window): one fc problem per position
                t0 = conv[k][1]
                m = k
                while m < j and conv[m][1].win == t0.win and conv[m][1].kind == t0.kind:
                    m += 1
                y0, y1, x0, x1 = t0.win
                khp, kwp = y1 - y0, x1 - x0
                P = khp * kwp
                cnt = m - k
                for q in range(k, m):
                    self.conv_rows.append((row + (q - k) * P, P, kwp, khp))
                    self.seg_of.append((seg_base + (row + (q - k) * P - seg_row0) * ld, ld))
                if t0.kind == KIND_CLS_W:
                    self.cls_heads.append((seg_base + (row - seg_row0) * ld, ld, ii, cnt, clsw_elems))
                    for q in range(k, m):
                        # cls nodes read the class-head output [classes][cnt * i'] (element (cls, b) of node j at
                        # cls * cnt*i' + j*i' + b): base offset + column stride of the class index
                        self.seg_of[q] = (clsw_elems + (q - k) * ii, cnt * ii)
                    clsw_elems += cnt * ncls * ii
                row += cnt * P
                k = m
            seg_rows = row - seg_row0
            self.segments.append((o, ii, seg_row0, seg_rows, seg_base))
            if ii == ms1:
                c2_probs.append((seg_row0, 0, seg_rows, o * ms1, seg_base, ld, 0))
            elif ii <= 128:
                c2_grouped.setdefault(ii, []).append((seg_row0, 0, seg_rows, o * ii, seg_base, ld, 0))
            else:
                for a in range(o):
                    c2_probs.append((seg_row0, a * ms1, seg_rows, ii, seg_base + a * ii, ld, a * ms1))
            wout_elems += seg_rows * ld
            i = j
        self.conv_total_rows = row
        self.wout_elems = wout_elems
        # ---- fc stage: ONE problem per decoder-grid position over all nodes whose crop window contains it.
        # The decoder-input rows (dec_in) are ordered by window, largest first (centred windows are nested, so the
        # nodes that need a position form a prefix / a few runs); the (node, position)-major h0 rows keep the
        # class-major order above through a row map.
        fc_order = sorted(range(len(conv)), key=lambda q: (-(conv[q][1].win[1] - conv[q][1].win[0]) *
                                                           (conv[q][1].win[3] - conv[q][1].win[2]),
                                                           conv[q][1].win, q))
        for r, q in enumerate(fc_order):
            b, t = conv[q]
            dst_row[offs[b] + t.node] = r
        wins = np.array([conv[q][1].win for q in fc_order], dtype=np.int64).reshape(-1, 4)
        rowmap = np.zeros(0, dtype=np.int64)
        if len(wins):
            # need[pos, r]: decoder-grid position pos = py * S + px lies inside the window of the r-th node (fc order);
            # one fc problem per run of consecutive r at a position, (node, position) rows through the row map
            py, px = np.divmod(np.arange(S * S, dtype=np.int64), S)
            need = ((wins[None, :, 0] <= py[:, None]) & (py[:, None] < wins[None, :, 1]) &
                    (wins[None, :, 2] <= px[:, None]) & (px[:, None] < wins[None, :, 3]))
            pos_i, r_i = np.nonzero(need)                              # position-major, r ascending
            first = np.ones(len(pos_i), dtype=bool)
            first[1:] = (pos_i[1:] != pos_i[:-1]) | (r_i[1:] != r_i[:-1] + 1)
            starts = np.nonzero(first)[0]
            lens = np.diff(np.append(starts, len(pos_i)))
            order_arr = np.asarray(fc_order, dtype=np.int64)
            rows_arr = np.asarray(self.conv_rows, dtype=np.int64).reshape(-1, 4)     # (row0, P, kwp, khp) per conv node
            q_i = order_arr[r_i]
            wq = wins[r_i]
            rowmap = rows_arr[q_i, 0] + (py[pos_i] - wq[:, 0]) * rows_arr[q_i, 2] + (px[pos_i] - wq[:, 2])
            for st_, ln_ in zip(starts.tolist(), lens.tolist()):
                pos = int(pos_i[st_])
                fc_probs.append((int(r_i[st_]), pos * 4 * C, ln_, 4 * C, st_, 4 * C, pos * 4 * C))
        self.fc_rowmap = np.asarray(rowmap, dtype=np.int32)
        self.clsw_elems = clsw_elems
        self.fc_problems = np.array(fc_probs, dtype=PROBLEM_DT) if fc_probs else np.zeros(0, PROBLEM_DT)
        self.c2_problems = np.array(c2_probs, dtype=PROBLEM_DT) if c2_probs else np.zeros(0, PROBLEM_DT)
        self.fc_tiles = tiles_for(self.fc_problems)
        # most grid positions are needed by a handful of nodes only: the "swapped" GEMM formulation (weights on the
        # 128 UMMA-M rows, 64 activation rows per tile) keeps all epilogue warps busy there
        self.fc_swap = len(self.fc_problems) > 0 and float(np.median(self.fc_problems['m'])) <= 64
        self.fc_tiles_swap = tiles_for(self.fc_problems, block_m=64) if self.fc_swap else self.fc_tiles
        self.c2_block_n = C2_DENSE_BLOCK_N
        self.c2_tiles = tiles_for(self.c2_problems, block_n=self.c2_block_n)
        # order conv2 tiles by weight block so CTAs that run together share the streamed weight rows through L2
        if len(self.c2_tiles):
            pr = self.c2_problems
            wrow = pr['b_row0'][self.c2_tiles[:, 0]].astype(np.int64) + \
                self.c2_tiles[:, 2].astype(np.int64) * self.c2_block_n
            self.c2_tiles = self.c2_tiles[np.lexsort((self.c2_tiles[:, 1], wrow))]
        # conv.2 launches: (b_group, problems, tiles); b_group 0 = dense rows, g > 0 = compact columns via 3-D boxes
        self.c2_launches = []
        if len(self.c2_problems):
            self.c2_launches.append((0, self.c2_problems, self.c2_tiles))
        for g_, probs in sorted(c2_grouped.items()):
            pr = np.array(probs, dtype=PROBLEM_DT)
            tl = tiles_for(pr, block_n=(128 // g_) * g_)
            tl = tl[np.lexsort((tl[:, 1], tl[:, 2]))]
            self.c2_launches.append((g_, pr, tl))

        # ---- 1-D nodes (decoder_1d), classification-bias nodes last ----
        one_d.sort(key=lambda bt: (bt[1].kind == KIND_CLS_B, bt[0], bt[1].node))
        self.one_d = one_d
        self.n_1d = len(one_d)
        self.n_clsb = sum(1 for _, t in one_d if t.kind == KIND_CLS_B)
        for r, (b, t) in enumerate(one_d):
            dst_row[offs[b] + t.node] = self.n_conv + r
        self.dst_row = dst_row
        self.max_ch = max_ch
        self._build_descs(offs, ncls, ms0, ms1, max_ch)

    # ------------------------------------------------------------------------------------------------------------
    def _build_descs(self, offs, ncls, ms0, ms1, max_ch):
        """One record per predicted tensor: static descriptor fields + (module, attr) + source buffer id/offset."""
        recs, targets, srcs = [], [], []
        self.n_tok_elems = 0

        cur_model = [0]

        FIELDS = ('t1', 't2', 't3', 'so', 'si', 'ld', 'ca', 'ra', 'kh_src', 'kw_src', 'cy', 'cx', 'scale', 'mode')

        def add(module, attr, shape, src_buf, src_off, t1=1, t2=1, t3=1, so=1, si=1, ld=0, ca=0, ra=0, kh_src=1,
                kw_src=1, cy=0, cx=0, scale=1.0, mode=0, view='full'):
            # plain tuples here, ONE structured array at the end (a numpy record per call costs ~40 us)
            numel = 1
            for v in shape:
                numel *= int(v)
            if numel >= (1 << 31):
                raise NotImplementedError('target tensors with 2^31 or more elements are not supported')
            recs.append((numel, cur_model[0], t1, t2, t3, so, si, ld, ca, ra, kh_src, kw_src, cy, cx, scale, mode))
            # view: 'full' = the whole parameter; 'tok' / 'body' = row 0 / rows 1.. of a ViT pos_embedding
            targets.append((module, attr, tuple(shape), view))
            srcs.append((src_buf, int(src_off)))

        for q, (b, t) in enumerate(self.conv):
            e = t.entry
            cur_model[0] = b
            mod, tsz = e['module'], e['sz']
            attr = param_attr(mod, e['is_w'])
            base, ld = self.seg_of[q]
            _, P, kwp, khp = self.conv_rows[q]
            if t.kind == KIND_CLS_W:
                tt = tuple(tsz) + (1,) * (4 - len(tsz))
                add(mod, attr, tsz, SRC_CLSW, base, t1=tt[1], t2=tt[2], t3=tt[3], so=ncls, si=t.i_need, ld=0,
                    ca=ld, scale=scale_for(tsz))
                continue
            common = dict(so=t.o_need, si=t.i_need, ld=ld, ca=t.i_need, kh_src=khp, kw_src=kwp)
            if len(tsz) == 3:
                # positional encoding (nn.py:441-446): row 0 = fresh N(0, 0.02) class token, rows 1.. = transposed map
                D = tsz[2]
                n_tok = min(tsz[1] - 1, P)
                if tsz[0] != 1 or n_tok != tsz[1] - 1:
                    raise NotImplementedError('positional encoding of shape %s is not supported' % (tsz,))
                add(mod, attr, (D,), SRC_TOK, self.n_tok_elems, so=t.i_need, ca=1, view='tok')
                self.n_tok_elems += t.i_need
                add(mod, attr, (n_tok, D), SRC_WOUT, base, t1=D, so=1 << 30, si=t.i_need, ld=ld, ca=0, ra=1,
                    view='body')
            elif len(tsz) == 4:
                if t.interp:
                    add(mod, attr, tsz, SRC_WOUT, base, t1=tsz[1], t2=tsz[2], t3=tsz[3], mode=3, scale=scale_for(tsz),
                        **common)
                else:
                    cy = khp // 2 - min(tsz[2], khp) // 2
                    cx = kwp // 2 - min(tsz[3], kwp) // 2
                    if tsz[2] > khp or tsz[3] > kwp:
                        raise NotImplementedError('target kernel %s larger than the decoded window' % (tsz,))
                    add(mod, attr, tsz, SRC_WOUT, base, t1=tsz[1], t2=tsz[2], t3=tsz[3], cy=cy, cx=cx,
                        scale=scale_for(tsz), **common)
            elif len(tsz) == 2:
                add(mod, attr, tsz, SRC_WOUT, base, t1=tsz[1], cy=khp // 2, cx=kwp // 2, scale=scale_for(tsz), **common)
            else:
                raise NotImplementedError('conv-decoded tensor of shape %s' % (tsz,))

        n_plain = self.n_1d - self.n_clsb
        for r, (b, t) in enumerate(self.one_d):
            e = t.entry
            cur_model[0] = b
            mod, tsz, is_w = e['module'], e['sz'], e['is_w']
            if t.kind == KIND_CLS_B:
                # bias_class output [(node, row)][classes], row 1 (nn.py:294,317)
                add(mod, param_attr(mod, False), tsz, SRC_CLSB, ((r - n_plain) * 2 + 1) * ncls, so=ncls, ca=1, mode=2)
                continue
            row_off = r * 2 * max_ch
            if t.kind == KIND_3D:
                # element (a, 0, c) = d1[a] * sqrt(1 / k): rows a, columns c (t1 = k), source column stride 0
                add(mod, param_attr(mod, is_w), tsz, SRC_D1, row_off, t1=int(tsz[1] * tsz[2]), so=2 * max_ch, si=1, ca=1,
                    scale=scale_for(tsz), mode=0)
                continue
            if is_w:
                add(mod, param_attr(mod, True), tsz, SRC_D1, row_off, so=max_ch, ca=1, mode=1)
                if getattr(mod, 'bias', None) is not None:
                    add(mod, param_attr(mod, False), tsz, SRC_D1, row_off + max_ch, so=max_ch, ca=1, mode=2)
            else:
                add(mod, param_attr(mod, False), tsz, SRC_D1, row_off + max_ch, so=max_ch, ca=1, mode=2)

        self.desc_static = np.zeros(len(recs), dtype=DESC_DT)
        if recs:
            cols = list(zip(*recs))
            self.desc_static['numel'], self.desc_static['norm_slot'] = cols[0], cols[1]
            for k_, col in zip(FIELDS, cols[2:]):
                self.desc_static[k_] = col
            for f_ in ('t1', 't2', 't3', 'so', 'si'):
                self.desc_static['m_' + f_], self.desc_static['s_' + f_] = fastdiv_array(self.desc_static[f_])
        self.desc_targets = targets
        self.desc_src_buf = np.array([s_[0] for s_ in srcs], dtype=np.int64)
        self.desc_src_off = np.array([s_[1] for s_ in srcs], dtype=np.uint64)
        # rows 1.. of a ViT pos_embedding start one row (D floats) after the parameter's base address
        def last_dim(p):
            return (p[-1] if isinstance(p, (list, tuple)) else p.shape[-1])
        self.desc_dst_shift = np.array([last_dim(getattr(tg[0], tg[1])) * 4 if tg[3] == 'body' else 0
                                        for tg in targets], dtype=np.uint64)
        chunks = (self.desc_static['numel'] + SCATTER_CHUNK - 1) // SCATTER_CHUNK
        self.desc_static['chunk0'] = np.concatenate([[0], np.cumsum(chunks)[:-1]]) if len(chunks) else []
        self.n_chunks = int(chunks.sum())
        self.chunk_desc = np.repeat(np.arange(len(chunks), dtype=np.int32), chunks)
