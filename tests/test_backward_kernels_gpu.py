"""Adjoint kernels of the training path (B200): each against torch autograd of the same fp32 op."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from ghn3_b200 import _lib as L
from ghn3_b200 import ops
from ghn3_b200.weights import CONFIGS
from tests import helpers as H
from tests.test_kernels_gpu import _pack, _desc_array

DEV = 'cuda'


def _rel(a, b):
    return H.max_rel_err(a.float(), b.float())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('shape', [(152, 384), (33, 70), (1, 16), (1000, 1536)])
def test_transpose(shape, dtype):
    torch.manual_seed(0)
    src = torch.randn(*shape, device=DEV).to(dtype)
    out = ops.transpose(src)
    assert out.shape == (shape[1], shape[0]) and out.stride(0) % 8 == 0
    assert torch.equal(out, src.t())
    out16 = ops.transpose(src, dst_dtype=ops.BF16)
    assert torch.equal(out16, src.t().bfloat16())


def test_transpose_grouped_rows():
    """rows taken as o' x i' sub-blocks of an ms x ms grid (decoder conv.2 weight rows, reference nn.py:750)."""
    torch.manual_seed(1)
    ms, g, outer, K = 24, 5, 7, 40
    w = torch.randn(ms * ms, K, device=DEV)
    out = ops.transpose(w, group=g, group_stride=ms, rows=outer * g)
    idx = torch.tensor([o * ms + i for o in range(outer) for i in range(g)], device=DEV)
    assert torch.equal(out, w[idx].t())


def test_elementwise_ops():
    torch.manual_seed(2)
    u = (torch.randn(777, 33, device=DEV) * 2).requires_grad_()
    dg = torch.randn(777, 33, device=DEV)
    g = F.gelu(u)
    g.backward(dg)
    assert _rel(ops.elementwise(L.EW_GELU, u.detach()), g.detach()) < 1e-6
    assert _rel(ops.elementwise(L.EW_GELU_BWD, dg, u.detach()), u.grad) < 1e-5
    y = torch.relu(torch.randn(50, 7, device=DEV))
    assert torch.equal(ops.elementwise(L.EW_RELU_BWD, dg[:50, :7].contiguous(), y), dg[:50, :7] * (y > 0))
    assert torch.equal(ops.elementwise(L.EW_ADD, dg, u.detach()), dg + u.detach())
    assert torch.equal(ops.elementwise(L.EW_COPY, dg, out_dtype=ops.BF16), dg.bfloat16())
    # bf16 operands
    ub = u.detach().bfloat16().float().requires_grad_()
    F.gelu(ub).backward(dg.bfloat16().float())
    assert _rel(ops.elementwise(L.EW_GELU_BWD, dg.bfloat16(), u.detach().bfloat16()).float(), ub.grad) < 1e-2


@pytest.mark.parametrize('rows,cols', [(1, 5), (300, 384), (1025, 77)])
def test_colsum(rows, cols):
    torch.manual_seed(3)
    src = torch.randn(rows, cols, device=DEV)
    dst = torch.ones(cols, device=DEV)
    ops.colsum(src, dst)
    assert _rel(dst, 1 + src.double().sum(0).float()) < 1e-5
    dst2 = torch.zeros(cols, device=DEV)
    ops.colsum(src.bfloat16(), dst2)
    assert _rel(dst2, src.bfloat16().double().sum(0).float()) < 1e-5


def test_colsum_grouped_columns():
    ms, g, outer = 24, 5, 7
    src = torch.randn(40, outer * g, device=DEV)
    dst = torch.zeros(ms * ms, device=DEV)
    ops.colsum(src, dst, group=g, group_stride=ms)
    ref = torch.zeros(ms * ms, device=DEV)
    idx = torch.tensor([o * ms + i for o in range(outer) for i in range(g)], device=DEV)
    ref[idx] = src.sum(0)
    assert _rel(dst, ref) < 1e-5


@pytest.mark.parametrize('C_', [32, 384, 1024])
@pytest.mark.parametrize('acc', [False, True])
def test_layernorm_bwd(C_, acc):
    torch.manual_seed(4)
    M = 301
    x = (torch.randn(M, C_, device=DEV) * 3 + 1).requires_grad_()
    gamma = torch.randn(C_, device=DEV).requires_grad_()
    beta = torch.randn(C_, device=DEV).requires_grad_()
    dy = torch.randn(M, C_, device=DEV)
    F.layer_norm(x, (C_,), gamma, beta, 1e-5).backward(dy)
    dx0 = torch.randn(M, C_, device=DEV)
    dx = dx0.clone()
    dg, db = torch.zeros(C_, device=DEV), torch.zeros(C_, device=DEV)
    ops.layernorm_bwd(x.detach(), gamma.detach(), dy, dx, dg, db, accumulate=acc)
    assert _rel(dx - (dx0 if acc else 0), x.grad) < 2e-4
    assert _rel(dg, gamma.grad) < 1e-4 and _rel(db, beta.grad) < 1e-4


def test_layernorm_bwd_row_gather():
    """adjoint of the final LayerNorm's row scatter: node r's gradient is row dst_row[r] of the decoder-input grad."""
    torch.manual_seed(5)
    M, C_ = 40, 64
    dst_row = torch.full((M,), -1, dtype=torch.int32)
    perm = torch.randperm(M)[:25]
    dst_row[perm] = torch.randperm(25, dtype=torch.int32)
    x = torch.randn(M, C_, device=DEV).requires_grad_()
    gamma = torch.randn(C_, device=DEV).requires_grad_()
    beta = torch.zeros(C_, device=DEV).requires_grad_()
    ddec = torch.randn(25, C_, device=DEV)
    y = F.layer_norm(x, (C_,), gamma, beta, 1e-5)
    dy = torch.zeros(M, C_, device=DEV)
    dy[perm.to(DEV)] = ddec[dst_row[perm].long().to(DEV)]
    y.backward(dy)
    dx = torch.zeros(M, C_, device=DEV)
    dg, db = torch.zeros(C_, device=DEV), torch.zeros(C_, device=DEV)
    ops.layernorm_bwd(x.detach(), gamma.detach(), ddec.bfloat16().float(), dx, dg, db, dy_row=dst_row.to(DEV))
    assert _rel(dx, x.grad) < 2e-2 and _rel(dg, gamma.grad) < 2e-2 and _rel(db, beta.grad) < 2e-2
    dx.zero_(); dg.zero_(); db.zero_()
    ops.layernorm_bwd(x.detach(), gamma.detach(), ddec, dx, dg, db, dy_row=dst_row.to(DEV))
    assert _rel(dx, x.grad) < 2e-4 and _rel(dg, gamma.grad) < 1e-4 and _rel(db, beta.grad) < 1e-4


@pytest.mark.parametrize('cfg_name,archs', [('ghn3tiny', 'small'), ('ghn3tm8', 'small'), ('ghn3lm8', 'small'),
                                            ('ghn3xlm16', 'small'), ('ghn3xlm16', 'large'), ('ghn3tiny', 'large')])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, 'bf16_mma'])
def test_attention_bwd(cfg_name, archs, dtype):
    mma = dtype == 'bf16_mma'                  # tensor-core kernels fed by the forward's softmax statistics
    if mma:
        dtype = torch.bfloat16
    cfg = CONFIGS[cfg_name]
    C_, H_ = cfg['hid'], cfg['heads']
    D = C_ // H_
    # 'large': graphs with more nodes than one shared-memory tile (256) next to a small one
    archs = ['resnet18', 'squeezenet1_0', 'mobilenet_v3_small'] if archs == 'small' else \
        ['swin_v2_t', 'resnet18', 'efficientnet_b3']
    recs, pack = _pack(archs)
    torch.manual_seed(6)
    N = pack.total_nodes
    qkv = torch.randn(N, 3 * C_, device=DEV).to(dtype)
    d_out = torch.randn(N, C_, device=DEV).to(dtype)
    lut = (torch.randn(H_, 51 * 51, device=DEV) * 0.5)
    lse2 = torch.zeros(H_, N, device=DEV) if mma else None
    out = ops.attention(qkv, pack, lut, C_, H_, dtype=ops.BF16 if dtype == torch.bfloat16 else ops.F32, lse2=lse2)
    d_lut = torch.zeros_like(lut)
    d_qkv = ops.attention_bwd(qkv, out, d_out, pack, lut, C_, H_, d_lut=d_lut, fwd_lse2=lse2)
    torch.cuda.synchronize()

    qf = qkv.float().clone().requires_grad_()
    lf = lut.clone().requires_grad_()
    off, outs = 0, []
    for g, rec in enumerate(recs):
        n = rec['n']
        A = pack.spd_matrix(g).long()
        bias = lf[:, (A * 51 + A.t()).reshape(-1)].view(H_, n, n)
        q, k, v = qf[off:off + n].view(n, 3, H_, D).permute(1, 2, 0, 3)
        attn = (q @ k.transpose(-2, -1)) * D ** -0.5 + bias
        outs.append((attn.softmax(-1) @ v).transpose(0, 1).reshape(n, C_))
        off += n
    torch.cat(outs).backward(d_out.float())
    tol = 3e-2 if dtype == torch.bfloat16 else 2e-4
    assert _rel(d_qkv, qf.grad) < tol
    assert _rel(d_lut, lf.grad) < tol


def test_node_features_bwd():
    torch.manual_seed(7)
    C_, N = 64, 200
    sizes = dict(embed_op=15, embed_ch=392, embed_sp=11, cent_in=101, cent_out=101, dist_embed=1001)
    idx = {k: torch.randint(0, v, (N,), dtype=torch.int32, device=DEV) for k, v in sizes.items()}
    shape_idx = torch.stack([torch.randint(0, 392, (N,)), torch.randint(0, 392, (N,)), torch.randint(0, 11, (N,)),
                             torch.randint(0, 11, (N,))], 1).int().to(DEV).contiguous()
    dx = torch.randn(N, C_, device=DEV)
    grads = {k: torch.zeros(v, C_ if k not in ('embed_ch', 'embed_sp') else C_ // 4, device=DEV)
             for k, v in sizes.items()}
    a = L.NodeFeaturesBwdArgs(total_nodes=N, hid=C_, op=L.ptr(idx['embed_op']), shape_idx=L.ptr(shape_idx),
                              deg_in=L.ptr(idx['cent_in']), deg_out=L.ptr(idx['cent_out']),
                              dist0=L.ptr(idx['dist_embed']), dx=L.ptr(dx), d_embed_op=L.ptr(grads['embed_op']),
                              d_embed_ch=L.ptr(grads['embed_ch']), d_embed_sp=L.ptr(grads['embed_sp']),
                              d_cent_in=L.ptr(grads['cent_in']), d_cent_out=L.ptr(grads['cent_out']),
                              d_dist_embed=L.ptr(grads['dist_embed']))
    L.call('node_features_bwd', a, L.current_stream())
    torch.cuda.synchronize()
    for k in ('embed_op', 'cent_in', 'cent_out', 'dist_embed'):
        ref = torch.zeros_like(grads[k]).index_add_(0, idx[k].long(), dx)
        assert _rel(grads[k], ref) < 1e-5, k
    Q = C_ // 4
    ref_ch = torch.zeros_like(grads['embed_ch'])
    ref_ch.index_add_(0, shape_idx[:, 0].long(), dx[:, :Q])
    ref_ch.index_add_(0, shape_idx[:, 1].long(), dx[:, Q:2 * Q])
    ref_sp = torch.zeros_like(grads['embed_sp'])
    ref_sp.index_add_(0, shape_idx[:, 2].long(), dx[:, 2 * Q:3 * Q])
    ref_sp.index_add_(0, shape_idx[:, 3].long(), dx[:, 3 * Q:])
    assert _rel(grads['embed_ch'], ref_ch) < 1e-5 and _rel(grads['embed_sp'], ref_sp) < 1e-5


@pytest.mark.parametrize('C_,H_', [(32, 8), (384, 16)])
def test_edge_lut_bwd(C_, H_):
    torch.manual_seed(8)
    V = 51
    # operands on a dyadic grid: the hidden pre-activations are exact in fp32 and at least 1/16 away from zero, so
    # the ReLU mask cannot differ between summation orders
    E = (torch.randint(-2, 3, (257, C_), device=DEV).float() * 0.5).requires_grad_()
    w1 = (torch.randint(-1, 2, (C_, 2 * C_), device=DEV).float() * 0.25).requires_grad_()
    b1 = (torch.randint(-8, 9, (C_,), device=DEV).float() * 0.125 + 0.0625).requires_grad_()
    w2 = (torch.randn(H_, C_, device=DEV) / C_ ** 0.5).requires_grad_()
    b2 = torch.randn(H_, device=DEV).requires_grad_()
    d_lut = torch.randn(H_, V * V, device=DEV)
    a = torch.arange(V, device=DEV).repeat_interleave(V) + 2
    b = torch.arange(V, device=DEV).repeat(V) + 2
    hid = torch.relu(torch.cat([E[a], E[b]], 1) @ w1.t() + b1)
    lut = (hid @ w2.t() + b2).t()
    got = ops.edge_lut(E.detach(), w1.detach(), b1.detach(), w2.detach(), b2.detach(), vmax=V - 1)
    assert _rel(got, lut.detach()) < 1e-4
    lut.backward(d_lut)
    grads = {'edge_embed': torch.zeros_like(E), 'w1': torch.zeros_like(w1), 'b1': torch.zeros_like(b1),
             'w2': torch.zeros_like(w2), 'b2': torch.zeros_like(b2)}
    ops.edge_lut_bwd(E.detach(), w1.detach(), b1.detach(), w2.detach(), d_lut, V - 1, grads)
    torch.cuda.synchronize()
    for k, ref in (('edge_embed', E.grad), ('w1', w1.grad), ('b1', b1.grad), ('w2', w2.grad), ('b2', b2.grad)):
        assert _rel(grads[k], ref) < 2e-4, k


def test_scatter_bwd_matches_autograd_of_forward():
    """tile (repeat + chop), centre window, 1-D squashes and the bilinear mode: the adjoint equals the gradient of
    sum(forward * R) taken by finite composition of torch ops."""
    torch.manual_seed(9)
    so, si, ld = 6, 5, 40
    kh, kw = 3, 3
    src = torch.randn(kh * kw, ld, device=DEV)
    ents = []
    # (a) conv weight 14 x 7 x 3 x 3 tiled from 6 x 5
    ents.append(dict(shape=(14, 7, 3, 3), t1=7, t2=3, t3=3, so=so, si=si, ld=ld, ca=si, kh_src=kh, kw_src=kw,
                     scale=0.37, mode=0))
    # (b) centre tap 1x1 of the same source, 9 x 5
    ents.append(dict(shape=(9, 5, 1, 1), t1=5, t2=1, t3=1, so=so, si=si, ld=ld, ca=si, kh_src=kh, kw_src=kw, cy=1, cx=1,
                     scale=1.5, mode=0))
    # (c) 1-D weight / bias squashes, 13 elements tiled from 6
    ents.append(dict(shape=(13,), t1=1, so=so, si=1, ld=ld, ca=1, mode=1))
    ents.append(dict(shape=(13,), t1=1, so=so, si=1, ld=ld, ca=1, mode=2))
    # (d) bilinear 5 x 5 from the 3 x 3 window
    ents.append(dict(shape=(4, 3, 5, 5), t1=3, t2=5, t3=5, so=so, si=si, ld=ld, ca=si, kh_src=kh, kw_src=kw,
                     scale=0.9, mode=3))
    dsts, descs, chunk, chunk_desc = [], [], 0, []
    for i, e in enumerate(ents):
        dst = torch.zeros(e['shape'], device=DEV)
        d = L.ScatterDesc(dst=dst.data_ptr(), src=src.data_ptr(), numel=dst.numel(), chunk0=chunk,
                          t1=e.get('t1', 1), t2=e.get('t2', 1), t3=e.get('t3', 1), so=e['so'], si=e.get('si', 1),
                          ld=e['ld'], ca=e['ca'], ra=0, kh_src=e.get('kh_src', 1), kw_src=e.get('kw_src', 1),
                          cy=e.get('cy', 0), cx=e.get('cx', 0), scale=e.get('scale', 1.0), mode=e['mode'], norm_slot=-1)
        L.fill_fastdiv(d)
        n_ch = (dst.numel() + L.SCATTER_CHUNK - 1) // L.SCATTER_CHUNK
        chunk_desc += [i] * n_ch
        chunk += n_ch
        descs.append(d)
        dsts.append(dst)
    ddev = _desc_array(descs)
    cd = torch.tensor(chunk_desc, dtype=torch.int32, device=DEV)
    ops.scatter(ddev, len(descs), chunk, cd)
    torch.cuda.synchronize()
    # linearity gives the reference: d<fwd(src), R>/dsrc via a numerical Jacobian-vector product of the forward
    R = [torch.randn_like(d) for d in dsts]
    d_src = torch.zeros_like(src)
    gp = torch.tensor([r.data_ptr() for r in R], dtype=torch.int64, device=DEV)
    sp = torch.tensor([d_src.data_ptr()] * len(R), dtype=torch.int64, device=DEV)
    a = L.ScatterBwdArgs(descs=L.ptr(ddev), n_descs=len(descs), n_chunks=chunk, chunk_desc=L.ptr(cd), grads=L.ptr(gp),
                         d_src=L.ptr(sp))
    L.call('scatter_bwd', a, L.current_stream())
    torch.cuda.synchronize()

    def fwd_loss(s):
        src.copy_(s)
        ops.scatter(ddev, len(descs), chunk, cd)
        torch.cuda.synchronize()
        return sum((d.double() * r.double()).sum() for d, r in zip(dsts, R)).item()
    base = src.clone()
    used = d_src.nonzero()
    assert len(used) > 50
    eps = 1e-2
    rng = np.random.default_rng(0)
    for j in rng.choice(len(used), 25, replace=False):
        r, c = used[j].tolist()
        p, m = base.clone(), base.clone()
        p[r, c] += eps
        m[r, c] -= eps
        num = (fwd_loss(p) - fwd_loss(m)) / (2 * eps)
        assert abs(num - d_src[r, c].item()) < 2e-2 * max(1.0, abs(num)), (r, c, num, d_src[r, c].item())
    src.copy_(base)
    # elements the forward never reads get no gradient
    assert float(d_src[:, so * si + 5:].abs().max()) == 0.0


@pytest.mark.parametrize('dtype', [ops.BF16, ops.TF32])
def test_gemm_sparse_k_blocks(dtype):
    """ghn3_gemm_args.kb_list / kb_off: every M tile visits only its listed K blocks; an empty list leaves D alone."""
    torch.manual_seed(10)
    bk = 64 if dtype == ops.BF16 else 32
    M, N, K = 3 * 128 + 40, 200, 16 * bk
    a = torch.zeros(M, K, device=DEV)
    lists = [[0, 3, 4], [], [15], [2, 7, 8, 9]]
    for mt, blocks in enumerate(lists):
        for kb in blocks:
            a[mt * 128:(mt + 1) * 128, kb * bk:(kb + 1) * bk] = torch.randn(min(128, M - mt * 128), bk, device=DEV)
    b = torch.randn(N, K, device=DEV) / K ** 0.5
    a_in, b_in = (a.bfloat16(), b.bfloat16()) if dtype == ops.BF16 else (ops.convert(a, ops.TF32), ops.convert(b, ops.TF32))
    kb_list = torch.tensor(sum(lists, []), dtype=torch.int32, device=DEV)
    kb_off = torch.tensor(np.concatenate([[0], np.cumsum([len(x) for x in lists])]), dtype=torch.int32, device=DEV)
    out = torch.full((M, N), 7.0, device=DEV)
    g = L.GemmArgs(a=a_in.data_ptr(), a_rows=M, lda=K, b=b_in.data_ptr(), b_rows=N, ldb=K, k=K, in_dtype=dtype,
                   d=out.data_ptr(), out_dtype=ops.F32, bias=None, act=ops.ACT_NONE, b_dynamic=1,
                   kb_list=kb_list.data_ptr(), kb_off=kb_off.data_ptr())
    g.single = L.GemmProblem(a_row0=0, b_row0=0, m=M, n=N, d_off=0, ldd=N, bias_off=-1)
    L.call('gemm', g, L.current_stream())
    torch.cuda.synchronize()
    ref = (a_in.double() @ b_in.double().t()).float()
    assert bool((out[128:256] == 7.0).all())                       # empty list: untouched
    for mt in (0, 2, 3):
        sl = slice(mt * 128, min((mt + 1) * 128, M))
        assert _rel(out[sl], ref[sl]) < 2e-5
