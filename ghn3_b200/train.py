"""
Training path of the B200-native GHN-3: `ghn(nets, graphs, keep_grads=True)` (reference ghn3/nn.py:186-349 with the
keep_grads branch of `_set_params`, nn.py:526-545) and the adjoint program that carries the target networks' loss
gradient back into the GHN parameters (what autograd does for the reference under ghn3/trainer.py:238-411).

The forward pass is the prediction program with a Graphormer stack that keeps its activations
(ghn3_graphormer_train_fwd) and a scatter into one flat buffer whose slices become the target networks' parameters
(tensors with a grad_fn). The backward pass is one hand-written kernel sequence (include/ghn3_b200.h, "Training
path"): scatter^T -> decoders^T -> Graphormer^T -> node features^T / edge-bias LUT^T. There is no autograd tape inside
the GHN and no CPU path.
"""
import ctypes as ct

import numpy as np
import torch

from . import _lib as L
from . import ops
from .plan import SRC_CLSB, SRC_CLSW, SRC_D1, SRC_WOUT

OPC = {'graphormer_bwd': 8, 'transpose': 9, 'elementwise': 10, 'colsum': 11, 'layernorm_bwd': 12,
       'attention_bwd': 13, 'scatter_bwd': 14, 'node_features_bwd': 15, 'edge_lut_bwd': 16, 'fc_bwd': 17,
       'relu_transpose_bwd': 18, 'expand_cols': 19, 'memset': 20, 'gemm': 3, 'gemm_simt': 4}


def _pad8(n):
    return (int(n) + 7) // 8 * 8


# ----------------------------------------------------------------------------------------------------------------
# transposed weight copies (operands of the dgrad GEMMs); rebuilt whenever the device weights are
def transposed_weights(ghn, w):
    wt = w.get('T')
    if wt is not None:
        return wt
    act = w['act']

    def T(t):
        """[cols, rows] view of a buffer whose row stride is padded to 8 elements; refreshed with the weights (the
        transposes are appended AFTER the conversions in the prebuilt refresh sequence)."""
        view = ops.transpose(t, dst_dtype=act)
        base = view._base if view._base is not None else view
        w['refresh_ops'].append(('transpose', L.TransposeArgs(
            src=t.data_ptr(), src_dtype=ops.BF16 if t.dtype == torch.bfloat16 else ops.F32, ld_src=t.stride(0),
            rows=t.shape[0], cols=t.shape[1], group=0, group_stride=0, dst=base.data_ptr(), dst_dtype=act,
            ld_dst=base.stride(0))))
        return view
    layers = (L.LayerWeightsT * ghn.layers)()
    keep = []
    for l, t in enumerate(w['layers_keep']):
        d = dict(w_qkv_t=T(t['w_qkv']), w_out_t=T(t['w_out']), w_ff1_t=T(t['w_ff1']), w_ff2_t=T(t['w_ff2']))
        keep.append(d)
        for k, v in d.items():
            setattr(layers[l], k, v.data_ptr())
    wt = {'layers': layers, 'keep': keep, 'c0_wT': T(w['c0_w']), 'c2_wT': T(w['c2_w']), 'd1_w0T': T(w['d1_w0']),
          'd1_w1T': T(w['d1_w1']), 'cls_wT': T(w['cls_w'])}
    w['T'] = wt
    return wt


def _conv2_k_blocks(segments, R, KF, ms1, bk, dev):
    """Sparse-K block lists (ghn3_gemm_args.kb_list / kb_off) of the two conv.2 backward GEMMs over the
    column-expanded operands: class (o', i') at rows [row0, row0 + rows) is non-zero only in the columns
    a * ms1 + b, a < o', b < i'.
      dgrad  dh1 = X . W2^T   : M tiles = 128 rows of X,  K blocks = `bk` expanded columns
      wgrad  dW2 = X^T . h1   : M tiles = 128 expanded columns, K blocks = `bk` rows"""
    n_cb, n_rt = -(-KF // bk), -(-R // 128)
    n_ct, n_rb = -(-KF // 128), -(-R // bk)
    d_mask = np.zeros((n_rt, n_cb), dtype=bool)            # dgrad: row tile x column block
    w_mask = np.zeros((n_ct, n_rb), dtype=bool)            # wgrad: column tile x row block
    for (o, ii, row0, rows, _) in segments:
        a = np.arange(o, dtype=np.int64) * ms1
        cols = np.zeros(KF + 1, dtype=np.int32)            # coverage of expanded columns via a difference array
        np.add.at(cols, a, 1)
        np.add.at(cols, a + ii, -1)
        used = np.cumsum(cols[:-1]) > 0
        cb = np.add.reduceat(used, np.arange(0, KF, bk)) > 0
        ct = np.add.reduceat(used, np.arange(0, KF, 128)) > 0
        rt0, rt1 = row0 // 128, (row0 + rows - 1) // 128
        rb0, rb1 = row0 // bk, (row0 + rows - 1) // bk
        d_mask[rt0:rt1 + 1] |= cb[None, :]
        w_mask[np.nonzero(ct)[0], rb0:rb1 + 1] = True

    def pack(mask):
        off = np.concatenate([[0], np.cumsum(mask.sum(1))]).astype(np.int32)
        lst = np.nonzero(mask)[1].astype(np.int32)
        if len(lst) == 0:
            lst = np.zeros(1, np.int32)
        return torch.from_numpy(lst).to(dev), torch.from_numpy(off).to(dev)
    return pack(d_mask), pack(w_mask)


FLAT_PAD = 1024


def flat_layout(ghn):
    """Order and offsets of the GHN parameters in the flat fp32 gradient buffer: [decoder, decoder_1d, bias_class |
    everything else], every tensor starting at a multiple of 4 elements. The first region is final as soon as the
    decoder adjoint has run, so its all-reduce (93% of the bytes at XL) overlaps the Graphormer adjoint.
    Returns (params in module order, params in buffer order, offsets [n+1], elements of the first region)."""
    params = list(ghn.parameters())
    early = {id(p) for m in (ghn.decoder, ghn.decoder_1d, ghn.bias_class) for p in m.parameters()}
    order = [p for p in params if id(p) in early] + [p for p in params if id(p) not in early]
    n_early = sum(1 for p in order if id(p) in early)
    # both regions are padded to a multiple of FLAT_PAD elements so that each splits into equal 16-byte aligned shards
    # for any world size dividing FLAT_PAD / 4 (reduce-scatter + sharded optimizer step, GradSync(shard=True))
    offs = np.zeros(len(order) + 1, dtype=np.int64)
    pos = 0
    for i, p in enumerate(order):
        if i == n_early:
            pos = (pos + FLAT_PAD - 1) // FLAT_PAD * FLAT_PAD
        offs[i] = pos
        pos += (p.numel() + 3) // 4 * 4
    early_elems = int(offs[n_early]) if n_early < len(order) else (pos + FLAT_PAD - 1) // FLAT_PAD * FLAT_PAD
    offs[-1] = (pos + FLAT_PAD - 1) // FLAT_PAD * FLAT_PAD
    return params, order, offs, early_elems


# ----------------------------------------------------------------------------------------------------------------
class _Backward:
    """The adjoint kernel sequence of one training program (built once per batch plan and weight version)."""

    def __init__(self, prog, ghn):
        self.prog, self.ghn = prog, ghn
        bp, w, dev = prog.bp, prog.w, prog.device
        act = w['act']
        adt = ops.TORCH_DTYPE[act]
        in_dt, x3 = w['dtype'], int(w['x3'])
        C, H = ghn.hid, ghn.heads
        ms0, ms1, S, _ = ghn.max_shape
        ncls, mc = ghn.num_classes, bp.max_ch
        N = bp.total_nodes
        wt = transposed_weights(ghn, w)
        self.keep = []
        self.ops = []
        self.zero = []                     # buffers cleared at the start of every backward pass
        Z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)
        E = lambda *shape, dtype=adt: torch.empty(*shape, dtype=dtype, device=dev)

        # ---- fp32 gradient of every GHN parameter: one flat buffer, views per parameter ----
        params, order, offs, self.early_elems = flat_layout(ghn)
        self.params = params
        # ONE flat gradient buffer per (GHN, device), shared by the backward programs of every batch plan (2.6 GB at
        # ghn3xlm16): meta-batches never repeat in real training, so a buffer per plan would grow without bound
        shared = ghn.__dict__.setdefault('_shared_gflat', {})
        gf = shared.get(dev)
        if gf is None or gf.numel() != int(offs[-1]):
            gf = shared[dev] = Z(int(offs[-1]))
        self.gflat = gf
        gof = {id(p): self.gflat[int(o):int(o) + p.numel()].view(p.shape) for o, p in zip(offs[:-1], order)}
        self.gviews = [gof[id(p)] for p in params]
        G = lambda p: gof[id(p)]
        self.zero.append(self.gflat)

        def add(name, args):
            self.ops.append((name, args))
            return args

        def transpose(src, rows, cols, ld_src, dst, src_dtype=None, group=0, group_stride=0):
            """dst [cols, pad8(rows)] <- src [rows(mapped), cols]; returns the padded row stride."""
            ld = _pad8(rows)
            add('transpose', L.TransposeArgs(src=src if isinstance(src, int) else src.data_ptr(),
                                             src_dtype=src_dtype, ld_src=ld_src, rows=rows, cols=cols, group=group,
                                             group_stride=group_stride, dst=dst.data_ptr(), dst_dtype=act, ld_dst=ld))
            return ld

        def ew(op, n, a, a_dt, b, b_dt, out, out_dt):
            P = lambda t: t if isinstance(t, int) or t is None else t.data_ptr()
            add('elementwise', L.ElementwiseArgs(op=op, n=n, a=P(a), a_dtype=a_dt, b=P(b), b_dtype=b_dt, out=P(out),
                                                 out_dtype=out_dt))

        def colsum(src, src_dt, rows, cols, ld, dst, group=0, group_stride=0):
            P = lambda t: t if isinstance(t, int) else t.data_ptr()
            add('colsum', L.ColsumArgs(src=P(src), src_dtype=src_dt, ld=ld, rows=rows, cols=cols, group=group,
                                       group_stride=group_stride, dst=P(dst)))

        def gemm(a, m, lda, b, n, ldb, k, d, out_dtype, accumulate=0, b_dynamic=1, rowmap=None, ldd=None, kb=None):
            P = lambda t: t if isinstance(t, int) else t.data_ptr()
            g = L.GemmArgs(a=P(a), a_rows=m, lda=lda, b=P(b), b_rows=n, ldb=ldb, k=k, in_dtype=in_dt, d=P(d),
                           out_dtype=out_dtype, bias=None, act=ops.ACT_NONE, accumulate=accumulate, tf32_x3=x3,
                           b_dynamic=b_dynamic, rowmap=L.ptr(rowmap))
            if kb is not None:
                g.kb_list, g.kb_off = kb[0].data_ptr(), kb[1].data_ptr()
            g.single = L.GemmProblem(a_row0=0, b_row0=0, m=m, n=n, d_off=0, ldd=n if ldd is None else ldd, bias_off=-1)
            add('gemm', g)

        el = 2 if act == ops.BF16 else 4
        F32 = ops.F32
        n_dec = bp.n_conv + bp.n_1d
        self.ddec = Z(max(n_dec, 1), C)
        self.zero.append(self.ddec)

        # ---- source-gradient buffers of the scatter ----
        dbufs = {}
        if bp.conv_total_rows > 0:
            self.dwout = Z(bp.wout_elems)
            dbufs[SRC_WOUT] = self.dwout
            self.zero.append(self.dwout)
            if bp.clsw_elems:
                self.dclsw = Z(bp.clsw_elems)
                dbufs[SRC_CLSW] = self.dclsw
                self.zero.append(self.dclsw)
        if bp.n_1d > 0:
            self.dd1 = Z(bp.n_1d, 2 * mc)
            dbufs[SRC_D1] = self.dd1
            self.zero.append(self.dd1)
            if bp.n_clsb:
                self.dclsb = Z(2 * bp.n_clsb, ncls)
                dbufs[SRC_CLSB] = self.dclsb
                self.zero.append(self.dclsb)
        n_desc = prog.n_desc
        self.n_desc = n_desc
        if n_desc:
            dsrc = np.zeros(n_desc, dtype=np.int64)
            for i in range(n_desc):
                b_ = int(bp.desc_src_buf[i])
                if b_ in dbufs:
                    dsrc[i] = dbufs[b_].data_ptr() + int(bp.desc_src_off[i]) * 4
            self.dsrc_dev = torch.from_numpy(dsrc).to(dev)
            self.grad_ptrs = torch.zeros(n_desc, dtype=torch.int64, device=dev)
            self.grad_ptrs_host = torch.zeros(n_desc, dtype=torch.int64).pin_memory()
            self.grad_ptrs_np = self.grad_ptrs_host.numpy()
            add('scatter_bwd', L.ScatterBwdArgs(descs=L.ptr(prog.desc_dev), n_descs=n_desc, n_chunks=bp.n_chunks,
                                                chunk_desc=L.ptr(prog.st['chunk_desc']), grads=L.ptr(self.grad_ptrs),
                                                d_src=L.ptr(self.dsrc_dev)))

        # ---- 1-D decoder (+ classification-bias head) ----
        dec1, bcl = ghn.decoder_1d, ghn.bias_class[1]
        if bp.n_1d > 0:
            n1 = bp.n_1d
            n1p = _pad8(n1)
            if bp.n_clsb:
                rows = 2 * bp.n_clsb
                d1blk = prog.d1.data_ptr() + (n1 - bp.n_clsb) * 2 * mc * 4
                dd1blk = self.dd1.data_ptr() + (n1 - bp.n_clsb) * 2 * mc * 4
                colsum(self.dclsb, F32, rows, ncls, ncls, G(bcl.bias))
                # dWbc[cls][k] = sum_row dY[row][cls] * relu(d1[row][k])
                add('gemm_simt', L.GemmSimtArgs(a=d1blk, sam=1, sak=mc, b=L.ptr(self.dclsb), sbn=1, sbk=ncls,
                                                bias=None, d=G(bcl.weight).data_ptr(), sdm=1, sdn=mc, m=mc, n=ncls,
                                                k=rows, relu_a=1, act=ops.ACT_NONE, batch=1))
                tmp = E(rows, mc, dtype=torch.float32)
                self.keep.append(tmp)
                add('gemm_simt', L.GemmSimtArgs(a=L.ptr(self.dclsb), sam=ncls, sak=1, b=L.ptr(w['bc_w']), sbn=1,
                                                sbk=mc, bias=None, d=L.ptr(tmp), sdm=mc, sdn=1, m=rows, n=mc, k=ncls,
                                                relu_a=0, act=ops.ACT_NONE, batch=1))
                ew(L.EW_RELU_BWD, rows * mc, tmp, F32, d1blk, F32, dd1blk, F32)
            dd1a = E(n1, 2 * mc)
            tA, tB = E(2 * max(mc, C), n1p), E(2 * C, n1p)
            dhid = E(n1, 2 * C)
            self.keep += [dd1a, tA, tB, dhid]
            d_in = prog.dec_in.data_ptr() + bp.n_conv * C * el
            ew(L.EW_COPY, n1 * 2 * mc, self.dd1, F32, None, 0, dd1a, act)
            colsum(self.dd1, F32, n1, 2 * mc, 2 * mc, G(dec1.fc[2].bias))
            transpose(self.dd1, n1, 2 * mc, 2 * mc, tA, src_dtype=F32)
            transpose(prog.hid1, n1, 2 * C, 2 * C, tB, src_dtype=act)
            gemm(tA, 2 * mc, n1p, tB, 2 * C, n1p, n1, G(dec1.fc[2].weight), F32, accumulate=1)
            gemm(dd1a, n1, 2 * mc, wt['d1_w1T'], 2 * C, wt['d1_w1T'].stride(0), 2 * mc, dhid, act, b_dynamic=0)
            ew(L.EW_RELU_BWD, n1 * 2 * C, dhid, act, prog.hid1, act, dhid, act)
            colsum(dhid, act, n1, 2 * C, 2 * C, G(dec1.fc[0].bias))
            transpose(dhid, n1, 2 * C, 2 * C, tA, src_dtype=act)
            transpose(d_in, n1, C, C, tB, src_dtype=act)
            gemm(tA, 2 * C, n1p, tB, C, n1p, n1, G(dec1.fc[0].weight), F32, accumulate=1)
            gemm(dhid, n1, 2 * C, wt['d1_w0T'], C, wt['d1_w0T'].stride(0), 2 * C,
                 self.ddec.data_ptr() + bp.n_conv * C * 4, F32, b_dynamic=0)

        # ---- conv decoder ----
        dec = ghn.decoder
        R = bp.conv_total_rows
        if R > 0:
            # class heads first: they add into dwout
            clp = dec.class_layer_predictor[1]
            for hi, (woff, ld, ii, cnt, coff) in enumerate(bp.cls_heads):
                n2 = cnt * ii
                n2p, nclsp = _pad8(n2), _pad8(ncls)
                tO, A2, rtT = E(n2, nclsp), E(ncls, n2p), E(ms0, n2p)
                drt = E(n2, ms0, dtype=torch.float32)
                self.keep += [tO, A2, rtT, drt]
                dout = self.dclsw.data_ptr() + coff * 4
                transpose(dout, ncls, n2, n2, tO, src_dtype=F32)                 # dOut^T  [n2][ncls]
                colsum(tO, act, n2, ncls, nclsp, G(clp.bias))
                transpose(tO, n2, ncls, nclsp, A2, src_dtype=act)                # dOut    [ncls][n2] (padded stride)
                transpose(prog.rts[hi], n2, ms0, ms0, rtT, src_dtype=act)        # rt^T    [ms0][n2]
                gemm(A2, ncls, n2p, rtT, ms0, n2p, n2, G(clp.weight), F32, accumulate=1)
                gemm(tO, n2, nclsp, wt['cls_wT'], ms0, wt['cls_wT'].stride(0), ncls, drt, F32, b_dynamic=0)
                add('relu_transpose_bwd', L.ReluTransposeBwdArgs(src=prog.wout.data_ptr() + woff * 4,
                                                                 d_src=self.dwout.data_ptr() + woff * 4, ld=ii,
                                                                 src_bs=ld, d_rt=L.ptr(drt), rows=ms0, cols=ii,
                                                                 batch=cnt))
            # conv.2: the per-class compact gradients are expanded into the full ms0*ms1 column space (zeros elsewhere,
            # written once: the expanded positions are the same every step), then ONE dgrad and ONE wgrad GEMM
            KF = ms0 * ms1
            rp = _pad8(R)
            self.X = torch.zeros(R, KF, dtype=adt, device=dev)
            self.XT = torch.zeros(KF, rp, dtype=adt, device=dev)
            segs, tile0 = (L.ExpandSeg * len(bp.segments))(), 0
            for i_, (o, ii, row0, rows, base) in enumerate(bp.segments):
                ld = o * ii
                tc = (ld + 31) // 32
                segs[i_] = L.ExpandSeg(src_off=base, row0=row0, rows=rows, ld=ld, group=0 if ii == ms1 else ii,
                                       tile0=tile0, tiles_c=tc)
                tile0 += ((rows + 31) // 32) * tc
            self.segs_dev = torch.from_numpy(np.frombuffer(bytes(segs), dtype=np.uint8).copy()).to(dev)
            c2w, c2b = dec.conv[2].weight, dec.conv[2].bias
            add('expand_cols', L.ExpandArgs(segs=L.ptr(self.segs_dev), n_segs=len(bp.segments), n_tiles=tile0,
                                            src=L.ptr(self.dwout), group_stride=ms1, x=L.ptr(self.X), ld_x=KF,
                                            xt=L.ptr(self.XT), ld_xt=rp, dtype=act, d_bias=G(c2b).data_ptr()))
            h1T = E(8 * C * rp)
            self.dh1, self.dh0 = E(R, 8 * C), E(R, 4 * C)
            self.keep.append(h1T)
            kb_d, kb_w = _conv2_k_blocks(bp.segments, R, KF, ms1, 64 if act == ops.BF16 else 32, dev)
            self.keep += [kb_d, kb_w]
            c2T = wt['c2_wT']
            gemm(self.X, R, KF, c2T, 8 * C, c2T.stride(0), KF, self.dh1, act, b_dynamic=0, kb=kb_d)
            transpose(prog.h1, R, 8 * C, 8 * C, h1T, src_dtype=act)
            gemm(self.XT, KF, rp, h1T, 8 * C, rp, R, G(c2w), F32, kb=kb_w)      # D pre-zeroed with the flat buffer
            rp = _pad8(R)
            h0T = E(4 * C * rp)
            self.keep.append(h0T)
            ew(L.EW_RELU_BWD, R * 8 * C, self.dh1, act, prog.h1, act, self.dh1, act)
            colsum(self.dh1, act, R, 8 * C, 8 * C, G(dec.conv[0].bias))
            transpose(self.dh1, R, 8 * C, 8 * C, h1T, src_dtype=act)
            transpose(prog.h0, R, 4 * C, 4 * C, h0T, src_dtype=act)
            gemm(h1T, 8 * C, rp, h0T, 4 * C, rp, R, G(dec.conv[0].weight), F32, accumulate=1)
            gemm(self.dh1, R, 8 * C, wt['c0_wT'], 4 * C, wt['c0_wT'].stride(0), 8 * C, self.dh0, act, b_dynamic=0)
            ew(L.EW_RELU_BWD, R * 4 * C, self.dh0, act, prog.h0, act, self.dh0, act)
            fcw = dec.fc[0].weight
            if not (fcw.is_contiguous() and fcw.dtype == torch.float32):
                raise RuntimeError('ghn3_b200: decoder.fc.0.weight must be a contiguous fp32 tensor for training')
            add('fc_bwd', L.FcBwdArgs(problems=L.ptr(prog.st['fc_problems']), n_problems=len(bp.fc_problems),
                                      max_m=int(bp.fc_problems['m'].max()), rowmap=L.ptr(prog.st['fc_rowmap']),
                                      dh0=L.ptr(self.dh0), dtype=act if act == ops.BF16 else F32,
                                      dec_in=L.ptr(prog.dec_in), fc_w=fcw.data_ptr(), hid=C, n_out=4 * C,
                                      grid_positions=S * S, d_fc_w=G(fcw).data_ptr(),
                                      d_fc_b=G(dec.fc[0].bias).data_ptr(), d_dec_in=L.ptr(self.ddec)))

        self.n_decoder_ops = len(self.ops)
        # ---- Graphormer stack ----
        mp = _pad8(N)
        self.lgr = (L.LayerGrads * ghn.layers)()
        for l, layer in enumerate(ghn.gnn):
            lg = self.lgr[l]
            lg.ln1_w, lg.ln1_b = G(layer.ln1.weight).data_ptr(), G(layer.ln1.bias).data_ptr()
            lg.w_qkv = G(layer.attn.to_qkv.weight).data_ptr()
            lg.w_out, lg.b_out = G(layer.attn.to_out[0].weight).data_ptr(), G(layer.attn.to_out[0].bias).data_ptr()
            lg.ln2_w, lg.ln2_b = G(layer.ln2.weight).data_ptr(), G(layer.ln2.bias).data_ptr()
            lg.w_ff1, lg.b_ff1 = G(layer.ff.net[0].weight).data_ptr(), G(layer.ff.net[0].bias).data_ptr()
            lg.w_ff2, lg.b_ff2 = G(layer.ff.net[3].weight).data_ptr(), G(layer.ff.net[3].bias).data_ptr()
        self.dx = E(N, C, dtype=torch.float32)
        ws = dict(dxa=E(N, C), dh=E(N, C), dhf=E(N, C, dtype=torch.float32), dqkv=E(N, 3 * C), dff=E(N, 4 * C),
                  ta=E(4 * C, mp), tb=E(4 * C, mp), lse=E(H, N, dtype=torch.float32),
                  delta=E(H, N, dtype=torch.float32))
        self.keep.append(ws)
        self.wt = wt
        self.d_lut = None                      # allocated when the pack (cutoff) is known
        self.gb = L.GraphormerBwdArgs(saved=ct.pointer(prog.ta), layers_t_host=wt['layers'], grads_host=self.lgr,
                                      d_ln_w=G(ghn.ln.weight).data_ptr() if ghn.layernorm else None,
                                      d_ln_b=G(ghn.ln.bias).data_ptr() if ghn.layernorm else None,
                                      d_dec_in=L.ptr(self.ddec), d_dec_dtype=F32, d_lut=None, dx=L.ptr(self.dx),
                                      m_pad=mp, **{k: L.ptr(v) for k, v in ws.items()})
        add('graphormer_bwd', self.gb)
        # ---- node features, edge-bias look-up table ----
        g0 = ghn.gnn[0]
        nf = prog.nf
        self.nfb = L.NodeFeaturesBwdArgs(total_nodes=N, hid=C, shape_idx=nf.shape_idx, dx=L.ptr(self.dx),
                                         d_embed_op=G(ghn.embed.weight).data_ptr(),
                                         d_embed_ch=G(ghn.shape_enc.embed_channel.weight).data_ptr(),
                                         d_embed_sp=G(ghn.shape_enc.embed_spatial.weight).data_ptr(),
                                         d_cent_in=G(g0.centrality_embed_in.weight).data_ptr(),
                                         d_cent_out=G(g0.centrality_embed_out.weight).data_ptr(),
                                         d_dist_embed=G(g0.input_dist_embed.weight).data_ptr())
        add('node_features_bwd', self.nfb)
        li = w['lut_inputs']
        self.lut_ws = None
        self.lb = L.EdgeLutBwdArgs(hid=C, heads=H, vmax=0, edge_embed=L.ptr(li[0]), w1=L.ptr(li[1]), b1=L.ptr(li[2]),
                                   w2=L.ptr(li[3]), d_lut=None, workspace=None,
                                   d_edge_embed=G(g0.attn.edge_embed.embed.weight).data_ptr(),
                                   d_w1=G(g0.attn.proj_e[0].weight).data_ptr(),
                                   d_b1=G(g0.attn.proj_e[0].bias).data_ptr(),
                                   d_w2=G(g0.attn.proj_e[2].weight).data_ptr(),
                                   d_b2=G(g0.attn.proj_e[2].bias).data_ptr())
        add('edge_lut_bwd', self.lb)
        # the buffers that accumulate atomically are cleared by memset ops at the head of the sequence
        head = [('memset', L.MemsetArgs(ptr=t.data_ptr(), bytes=t.numel() * t.element_size())) for t in self.zero]
        self.ops = head + self.ops
        self.n_decoder_ops += len(head)
        self.seq = (L.SeqOp * len(self.ops))()
        for i, (name, args) in enumerate(self.ops):
            self.seq[i].op = OPC[name]
            self.seq[i].args = ct.cast(ct.pointer(args), ct.c_void_p)

    # ------------------------------------------------------------------------------------------------------------
    def run(self, out_grads, out_index, flat_grad=None, slices=None):
        """out_grads: gradients of the program outputs (one per predicted PARAMETER, None allowed);
        out_index[i] = (output index, byte shift, has_grad_path) of descriptor i; flat_grad: gradient of the flat
        buffer all outputs are slices of (slices[oi] = (offset, numel, shape)). Returns the GHN parameter grads."""
        prog = self.prog
        pack = prog.bound_pack
        dev = prog.device
        vmax = pack.cutoff
        lut_size = (vmax + 1) ** 2
        if prog.lse2 is not None:                  # tensor-core attention backward: dS accumulation scratch
            need = int(pack.mat_off[-1]) * self.ghn.heads
            if getattr(self, 'ds_total', None) is None or self.ds_total.numel() < need:
                self.ds_total = torch.empty(max(need, 1), dtype=torch.float32, device=dev)
            self.gb.ds_total, self.gb.ds_total_bytes = self.ds_total.data_ptr(), need * 4
        if self.d_lut is None or self.d_lut.shape[1] != lut_size:
            self.d_lut = torch.zeros(self.ghn.heads, lut_size, dtype=torch.float32, device=dev)
            self.lut_ws = torch.empty(4 * (vmax + 1) * self.ghn.hid, dtype=torch.float32, device=dev)
            self.gb.d_lut = self.d_lut.data_ptr()
            self.lb.d_lut, self.lb.workspace, self.lb.vmax = self.d_lut.data_ptr(), self.lut_ws.data_ptr(), vmax
        nfb, nf = self.nfb, prog.nf
        nfb.op, nfb.deg_in, nfb.deg_out, nfb.dist0 = nf.op, nf.deg_in, nf.deg_out, nf.dist0
        self.d_lut.zero_()
        live = []
        if self.n_desc:
            # the pinned table is read by an asynchronous H2D copy: the previous step's copy must have run before the
            # host rewrites it (the Trainer step has no other host synchronisation)
            ev_ptrs = getattr(self, 'ev_ptrs', None)
            if ev_ptrs is not None:
                ev_ptrs.synchronize()
            host = self.grad_ptrs_np
            flat_ptr = 0
            if flat_grad is not None:
                if flat_grad.dtype != torch.float32 or not flat_grad.is_contiguous():
                    flat_grad = flat_grad.float().contiguous()
                live.append(flat_grad)
                flat_ptr = flat_grad.data_ptr()
            cache = {}
            for i, (oi, shift, has_path) in enumerate(out_index):
                if not has_path:
                    host[i] = 0
                    continue
                g = cache.get(oi, False)
                if g is False:
                    g = out_grads[oi]
                    if g is not None and flat_ptr:                # both routes carry gradient: add them
                        o, n, shape = slices[oi]
                        g = g + flat_grad[o:o + n].view(shape)
                    if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                        g = g.float().contiguous()
                    cache[oi] = g
                    if g is not None:
                        live.append(g)
                if g is not None:
                    host[i] = g.data_ptr() + shift
                elif flat_ptr:
                    host[i] = flat_ptr + slices[oi][0] * 4 + shift
                else:
                    host[i] = 0
            self.grad_ptrs.copy_(self.grad_ptrs_host, non_blocking=True)
            self.ev_ptrs = torch.cuda.Event()
            self.ev_ptrs.record()
        stream = L.current_stream()
        lib = L.load()
        nd = self.n_decoder_ops
        sync = getattr(self.ghn, '_grad_sync', None)
        prof = getattr(self.ghn, '_profile_bwd', None)
        if prof is not None:                       # per-op device times (tools/profile_train.py); no exchange
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            prof.append(('start', ev))
            for i, (name, args) in enumerate(self.ops):
                one = ct.c_void_p(ct.addressof(self.seq) + i * ct.sizeof(L.SeqOp))
                L.check(lib.ghn3_run_sequence(one, 1, ct.c_void_p(stream)), 'ghn3_run_sequence (%s)' % name)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                prof.append((name, ev))
            self.live = live
            return self.gviews
        L.check(lib.ghn3_run_sequence(self.seq, nd, ct.c_void_p(stream)), 'ghn3_run_sequence (decoder backward)')
        works = []
        if sync is not None and self.early_elems:
            works.append(sync.start(self.gflat[:self.early_elems]))     # overlaps the Graphormer adjoint below
        rest = ct.c_void_p(ct.addressof(self.seq) + nd * ct.sizeof(L.SeqOp))
        L.check(lib.ghn3_run_sequence(rest, len(self.ops) - nd, ct.c_void_p(stream)),
                'ghn3_run_sequence (Graphormer backward)')
        if sync is not None:
            works.append(sync.start(self.gflat[self.early_elems:]))
            sync.finish(works, self.gflat)
        self.live = live
        return self.gviews


# ----------------------------------------------------------------------------------------------------------------
class GradSync:
    """
    Data-parallel gradient exchange of the training path (what DistributedDataParallel does for the reference,
    ghn3/trainer.py:134-136): the mean over ranks of the flat fp32 gradient buffer, as two asynchronous collectives
    (decoder region while the Graphormer adjoint still runs, then the rest). NCCL averages in the collective; other
    backends (gloo in the CPU tests) sum and divide.

    shard=True (NCCL): each region is REDUCE-SCATTERED in place -- rank r ends up with the averaged gradient of the
    r-th equal slice of each region only -- for an optimizer that updates just that slice and all-gathers the
    parameters (FusedAdamW.enable_sharding): the exchange moves the same bytes as an all-reduce, but the optimizer pass
    over 2.6 GB of parameters / moments shrinks with the number of ranks.
    """

    def __init__(self, group=None, shard=False):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.avg = dist.get_backend(group) == 'nccl'
        self.shard = bool(shard) and self.avg and self.world > 1 and FLAT_PAD % (4 * self.world) == 0

    def shard_of(self, lo, hi):
        """This rank's slice [a, b) of the flat region [lo, hi) (hi - lo is a multiple of FLAT_PAD)."""
        n = (hi - lo) // self.world
        return lo + self.rank * n, lo + (self.rank + 1) * n

    def start(self, flat):
        op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
        if self.shard and flat.numel() > 0:
            a, b = self.shard_of(0, flat.numel())
            return self.dist.reduce_scatter_tensor(flat[a:b], flat, op=op, group=self.group, async_op=True)
        return self.dist.all_reduce(flat, op=op, group=self.group, async_op=True)

    def finish(self, works, flat):
        for w_ in works:
            w_.wait()
        if not self.avg and self.world > 1:
            flat.div_(self.world)


def enable_grad_sync(ghn, group=None, shard=False):
    """Makes every backward pass through `ghn` average the GHN gradients over the ranks of `group` (shard=True: each
    rank keeps only its slice of the average, see GradSync)."""
    ghn._grad_sync = GradSync(group, shard=shard)
    return ghn


# ----------------------------------------------------------------------------------------------------------------
class _PredictFn(torch.autograd.Function):
    """(GHN parameters) -> (predicted target parameters): forward = the prediction program, backward = _Backward."""

    @staticmethod
    def forward(ctx, ghn, prog, pack, out_meta, out_index, *params):
        ctx.set_materialize_grads(False)          # unused outputs arrive as None, not as zero tensors
        prog.bind_pack(pack)
        total = out_meta['total']
        # One flat buffer per program, reused by every step (like the saved activations): its address is stable, so
        # the scatter descriptors are uploaded once -- a per-step pageable H2D copy would serialise host and device.
        # The padding between slices is zeroed once and never written.
        pred = getattr(prog, 'pred_buf', None)
        if pred is None or pred.numel() != total or prog.pred_wn != bool(ghn.weight_norm):
            pred = torch.zeros(total, dtype=torch.float32, device=prog.device)
            prog.pred_buf, prog.pred_wn = pred, bool(ghn.weight_norm)
            prog.point_descriptors_at(pred, out_meta, ghn.weight_norm)
        if prog.bp.n_tok_elems:
            prog.tok.normal_(mean=0.0, std=0.02)
        prog.run(getattr(ghn, '_profile', None))
        prog.step_id = getattr(prog, 'step_id', 0) + 1
        ctx.ghn, ctx.prog, ctx.out_index, ctx.step_id = ghn, prog, out_index, prog.step_id
        ctx.slices = out_meta['slices']
        outs = tuple(pred.as_strided(shape, st, o) for (o, n, shape), st in zip(out_meta['slices'], out_meta['strides']))
        return outs + (pred,)

    @staticmethod
    def backward(ctx, *grads):
        prog, ghn = ctx.prog, ctx.ghn
        if prog.step_id != ctx.step_id:
            raise RuntimeError('ghn3_b200: backward() called after another forward pass of the same (GHN, batch) '
                               'program; its saved activations have been overwritten')
        if prog.bwd is None:
            prog.bwd = _Backward(prog, ghn)
        bwd = prog.bwd
        # ghn.direct_grads (set by Trainer, whose step is loss.backward() after zero_grad(set_to_none=True)): the
        # parameters' .grad become views of the backward pass's flat buffer directly. Returning them through autograd
        # instead makes AccumulateGrad clone all 2.6 GB (the views are shared) and the fused optimizer gather them
        # back into a flat buffer. Off by default because torch.autograd.grad() needs the returned values.
        direct = bool(getattr(ghn, 'direct_grads', False)) and all(p.grad is None for p in bwd.params)
        if not direct:
            lo = bwd.gflat.data_ptr()
            hi = lo + bwd.gflat.numel() * 4
            for p in bwd.params:                # accumulated gradients must not live in the buffer we are about to clear
                if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                    p.grad = p.grad.clone()
        gviews = bwd.run(grads[:-1], ctx.out_index, grads[-1], ctx.slices)
        if direct:
            for p, v in zip(bwd.params, gviews):
                p.grad = v
            return (None,) * (5 + len(gviews))
        return (None, None, None, None, None) + tuple(gviews)


def forward_keep_grads(ghn, nets, graphs, w, bp, return_embeddings):
    """GHN3.forward(keep_grads=True): predicted parameters are tensors with a grad_fn, set on the target modules the
    way the reference does (ghn3/nn.py:526-545)."""
    from .nn import _Program
    device = ghn.embed.weight.device
    pdl = getattr(ghn, 'programmatic_launch', None)       # one chain at a time here: programmatic launch pays
    L.set_programmatic_launch(True if pdl is None else pdl)
    L.set_persistent_ctas(getattr(ghn, 'persistent_ctas', None) or 0)      # one chain at a time: every SM
    prog = getattr(bp, 'train_program', None)
    if prog is None or prog.w is not w or prog.device != device or prog.want_emb != bool(return_embeddings):
        prog = _Program(ghn, w, bp, device, bool(return_embeddings), train=True)
        bp.train_program = prog
    meta = getattr(bp, 'out_meta', None)
    if meta is None:
        # one output per predicted PARAMETER (a ViT pos_embedding is written by two descriptors)
        index, slices, keys, off = [], [], {}, 0
        for (module, attr, shape, view) in bp.desc_targets:
            key = (id(module), attr)
            if key not in keys:
                p = getattr(module, attr)
                full = tuple(p) if isinstance(p, (list, tuple)) else tuple(p.shape)
                n = int(np.prod(full))
                keys[key] = len(slices)
                slices.append((off, n, full))
                off += (n + 3) // 4 * 4
            index.append(keys[key])
        def contiguous_strides(shape):
            st, acc = [], 1
            for d in reversed(shape):
                st.append(acc)
                acc *= d
            return tuple(reversed(st))
        meta = {'slices': slices, 'strides': [contiguous_strides(sh) for (_, _, sh) in slices], 'total': max(off, 4),
                'desc_out': index,
                'targets': [(m, a) for (m, a, _, v) in bp.desc_targets]}
        shifts = [int(s) for s in bp.desc_dst_shift]
        meta['out_index'] = [(index[i], shifts[i], bp.desc_targets[i][3] != 'tok') for i in range(len(index))]
        bp.out_meta = meta
    plist = ghn.__dict__.get('_param_list') or list(ghn.parameters())
    outs = _PredictFn.apply(ghn, prog, graphs.pack, meta, meta['out_index'], *plist)
    # assignment as in nn.py:526-545; the (module dict, parameter dict, attribute, output) table is built once
    assign = meta.get('assign')
    if assign is None:
        assign, done = [], set()
        for (module, attr), oi in zip(meta['targets'], meta['desc_out']):
            key = (id(module), attr)
            if key in done:
                continue
            done.add(key)
            cur = module.__dict__.get(attr, module._parameters.get(attr) if hasattr(module, '_parameters') else None)
            light = isinstance(cur, (list, tuple))         # light modules keep shapes, not parameters (nn.py:527-533)
            assign.append((module if light else module.__dict__, None if light else module._parameters, attr, oi))
        meta['assign'] = assign
    for d, params, attr, oi in assign:
        t = outs[oi]
        if params is None:
            setattr(d, attr, t)
        else:
            d[attr] = t                                    # nn.py:536-539
            params[attr] = t
    ghn.last_program = prog
    # every predicted parameter is a slice of this flat tensor (same autograd node): a loss written on it, e.g. the
    # <p, R> stub of the GHN-only training benchmark, costs one kernel instead of one per parameter
    prog.pred_flat = outs[-1]
    return prog.emb


# ----------------------------------------------------------------------------------------------------------------
class _PredWd(torch.autograd.Function):
    """predparam_wd * sum_p ||p||_F over the predicted parameters (reference ghn3/trainer.py:288-294) evaluated on the
    flat buffer that holds them all: two kernel passes forward (ghn3_segnorm mode 0), one backward (mode 1), instead
    of a torch.norm + its backward per tensor."""

    @staticmethod
    def forward(ctx, pred_flat, tables, coef):
        dev = pred_flat.device
        out = torch.empty((), dtype=torch.float32, device=dev)
        a = L.SegnormArgs(src=pred_flat.data_ptr(), seg_off=tables['off'].data_ptr(),
                          seg_numel=tables['numel'].data_ptr(), chunk0=tables['chunk0'].data_ptr(),
                          chunk_seg=tables['chunk_seg'].data_ptr(), n_chunks=tables['n_chunks'],
                          n_segs=tables['n_segs'], mode=0, sumsq=tables['sumsq'].data_ptr(), total=out.data_ptr(),
                          coef=float(coef))
        L.call('segnorm', a, L.current_stream())
        ctx.tables, ctx.coef = tables, float(coef)
        ctx.save_for_backward(pred_flat)
        return out

    @staticmethod
    def backward(ctx, g):
        pred_flat, = ctx.saved_tensors
        t = ctx.tables
        grad = torch.zeros_like(pred_flat)
        gs = g.detach().float().contiguous()
        a = L.SegnormArgs(src=pred_flat.data_ptr(), seg_off=t['off'].data_ptr(), seg_numel=t['numel'].data_ptr(),
                          chunk0=t['chunk0'].data_ptr(), chunk_seg=t['chunk_seg'].data_ptr(), n_chunks=t['n_chunks'],
                          n_segs=t['n_segs'], mode=1, sumsq=t['sumsq'].data_ptr(), grad=grad.data_ptr(),
                          gscale=gs.data_ptr(), coef=ctx.coef)
        L.call('segnorm', a, L.current_stream())
        ctx.keep = gs
        return grad, None, None


def predicted_param_decay(ghn, coef):
    """coef * sum over the predicted parameter tensors of the last keep_grads forward of their Frobenius norms, as a
    differentiable device scalar."""
    prog = ghn.last_program
    meta = prog.bp.out_meta
    tables = meta.get('segnorm')
    if tables is None or tables['device'] != prog.device:
        CH = 8192
        offs = np.array([o for (o, n, _) in meta['slices']], dtype=np.int64)
        numels = np.array([n for (o, n, _) in meta['slices']], dtype=np.int64)
        chunks = (numels + CH - 1) // CH
        chunk0 = np.concatenate([[0], np.cumsum(chunks)[:-1]]).astype(np.int64)
        to_dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(prog.device)
        tables = {'off': to_dev(offs), 'numel': to_dev(numels), 'chunk0': to_dev(chunk0),
                  'chunk_seg': to_dev(np.repeat(np.arange(len(offs), dtype=np.int32), chunks)),
                  'n_chunks': int(chunks.sum()), 'n_segs': len(offs), 'device': prog.device,
                  'sumsq': torch.zeros(len(offs), dtype=torch.float64, device=prog.device)}
        meta['segnorm'] = tables
    return _PredWd.apply(prog.pred_flat, tables, coef)
