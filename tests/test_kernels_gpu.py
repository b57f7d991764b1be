"""Per-kernel parity tests (B200): every C-ABI entry point against the oracle / a plain PyTorch fp32 reference."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ghn3_b200 import _lib as L
from ghn3_b200 import ops
from ghn3_b200.weights import CONFIGS, procedural_state_dict
from oracle import ghn3_oracle as O
from tests import helpers as H

DEV = 'cuda'


def _rel(a, b):
    return H.max_rel_err(a.float(), b.float())


# --------------------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(152, 1152, 384), (128, 128, 64), (300, 1536, 1536), (5, 40, 32), (81, 1000, 3072), (1, 384, 384),
               (816, 384, 1536), (257, 200, 72)]


@pytest.mark.parametrize('m,n,k', GEMM_SHAPES)
@pytest.mark.parametrize('dtype', [ops.BF16, ops.TF32])
def test_gemm_plain(m, n, k, dtype):
    torch.manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    if dtype == ops.BF16:
        a_in, b_in = a.bfloat16(), b.bfloat16()
        a_ref, b_ref = a_in.float(), b_in.float()
    else:
        a_in, b_in = ops.convert(a, ops.TF32), ops.convert(b, ops.TF32)
        a_ref, b_ref = a_in, b_in
    ref = (a_ref.double() @ b_ref.double().t()).float()
    out = ops.gemm(a_in, b_in, in_dtype=dtype, out_dtype=ops.F32)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-5, (m, n, k, _rel(out, ref))


@pytest.mark.parametrize('m,n,k', GEMM_SHAPES)
def test_gemm_tf32_x3_is_fp32_accurate(m, n, k):
    """3-term compensated tf32: un-rounded fp32 operands, error ~1e-6 instead of ~5e-4."""
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    ref = (a.double() @ b.double().t()).float()
    out = ops.gemm(a, b, in_dtype=ops.TF32, out_dtype=ops.F32, x3=True)
    out1 = ops.gemm(ops.convert(a, ops.TF32), ops.convert(b, ops.TF32), in_dtype=ops.TF32, out_dtype=ops.F32)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 3e-5, (m, n, k, _rel(out, ref))
    assert _rel(out1, ref) < 3e-3


@pytest.mark.parametrize('x3', [False, True])
@pytest.mark.parametrize('splits', [0, 2, 5])
def test_gemm_split_k_accumulate(splits, x3):
    torch.manual_seed(splits)
    m, n, k = 457, 384, 1536
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    bias = torch.randn(n, device=DEV)
    a_in, b_in, dt = (a, b, ops.TF32) if x3 else (a.bfloat16(), b.bfloat16(), ops.BF16)
    x = torch.randn(m, n, device=DEV)
    ref = (x.double() + a_in.double() @ b_in.double().t() + bias.double()).float()
    ops.gemm(a_in, b_in, bias=bias, in_dtype=dt, out=x, out_dtype=ops.F32, accumulate=True, k_splits=splits, x3=x3)
    torch.cuda.synchronize()
    assert _rel(x, ref) < 2e-5


@pytest.mark.parametrize('block_n', [64, 128, 256])
def test_gemm_block_n(block_n):
    torch.manual_seed(block_n)
    a = torch.randn(200, 512, device=DEV).bfloat16()
    b = (torch.randn(700, 512, device=DEV) / 22).bfloat16()
    ref = (a.double() @ b.double().t()).float()
    out = ops.gemm(a, b, in_dtype=ops.BF16, out_dtype=ops.F32, block_n=block_n)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize('mode', ['bf16', 'tf32', 'x3'])
def test_gemm_large_single_problem_persistent(mode):
    """One problem with >= 2 x #SMs tiles on the persistent kernel with enumerated tiles (ghn3_gemm_args.persistent_single;
    the Graphormer GEMMs of an all-architecture batch: M = 18 666, K = C): bias + GELU into the activation dtype, residual accumulation, ragged
    M / N edges."""
    torch.manual_seed(5)
    m, n, k = 128 * 40 + 37, 128 * 9 + 72, 384            # 41 x 10 = 410 tiles
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    bias = torch.randn(n, device=DEV)
    if mode == 'bf16':
        a_in, b_in, dt, x3, tol = a.bfloat16(), b.bfloat16(), ops.BF16, False, 2e-5
    elif mode == 'tf32':
        a_in, b_in, dt, x3, tol = ops.convert(a, ops.TF32), ops.convert(b, ops.TF32), ops.TF32, False, 2e-5
    else:
        a_in, b_in, dt, x3, tol = a, b, ops.TF32, True, 3e-5
    lin = a_in.double() @ b_in.double().t() + bias.double()
    out = ops.gemm(a_in, b_in, bias=bias, in_dtype=dt, out_dtype=ops.F32, x3=x3, block_n=128, persistent_single=True)
    torch.cuda.synchronize()
    assert _rel(out, lin.float()) < tol
    act_dt = ops.BF16 if mode == 'bf16' else ops.F32
    out = ops.gemm(a_in, b_in, bias=bias, act=ops.ACT_GELU, in_dtype=dt, out_dtype=act_dt, x3=x3, block_n=128,
                   persistent_single=True)
    torch.cuda.synchronize()
    assert _rel(out, torch.nn.functional.gelu(lin).float()) < (1e-2 if mode == 'bf16' else 1e-4)
    x = torch.randn(m, n, device=DEV)
    ref = (x.double() + lin).float()
    ops.gemm(a_in, b_in, bias=bias, in_dtype=dt, out=x, out_dtype=ops.F32, accumulate=True, x3=x3, block_n=128,
             persistent_single=True)
    torch.cuda.synchronize()
    assert _rel(x, ref) < tol


@pytest.mark.parametrize('dtype', [ops.BF16, ops.TF32])
def test_gemm_epilogues(dtype):
    torch.manual_seed(1)
    m, n, k = 305, 1536, 384
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    bias = torch.randn(n, device=DEV)
    if dtype == ops.BF16:
        a_in, b_in = a.bfloat16(), b.bfloat16()
    else:
        a_in, b_in = ops.convert(a, ops.TF32), ops.convert(b, ops.TF32)
    lin = (a_in.double() @ b_in.double().t() + bias.double())
    # bias + GELU, activation-dtype output
    out = ops.gemm(a_in, b_in, bias=bias, act=ops.ACT_GELU, in_dtype=dtype, out_dtype=dtype)
    ref = torch.nn.functional.gelu(lin).float()
    torch.cuda.synchronize()
    assert _rel(out, ref) < (6e-3 if dtype == ops.BF16 else 6e-4)
    # bias + ReLU fp32
    out = ops.gemm(a_in, b_in, bias=bias, act=ops.ACT_RELU, in_dtype=dtype, out_dtype=ops.F32)
    torch.cuda.synchronize()
    assert _rel(out, torch.relu(lin).float()) < 2e-5
    # residual accumulate in place
    x = torch.randn(m, n, device=DEV)
    x0 = x.clone()
    ops.gemm(a_in, b_in, bias=bias, in_dtype=dtype, out=x, out_dtype=ops.F32, accumulate=True)
    torch.cuda.synchronize()
    assert _rel(x, (x0.double() + lin).float()) < 2e-5


def test_gemm_grouped_strided_output():
    """Grouped launch: problems with row offsets into A and B, strided D (the decoder's fc / conv2 pattern)."""
    torch.manual_seed(2)
    K = 384
    a = torch.randn(40, K, device=DEV).bfloat16()
    b = (torch.randn(2048, K, device=DEV) / 20).bfloat16()
    bias = torch.randn(2048, device=DEV)
    # problem 0: rows 0..9 of A, B rows 256..639 -> D0 [10, 384] at offset 0, ld 500
    # problem 1: rows 10..39 of A, B rows 1024..1223 (n=200) -> D1 at offset 10*500, ld 500, no bias
    probs = np.zeros(2, dtype=[('a_row0', 'i4'), ('b_row0', 'i4'), ('m', 'i4'), ('n', 'i4'), ('d_off', 'i8'),
                               ('ldd', 'i4'), ('bias_off', 'i4')])
    probs[0] = (0, 256, 10, 384, 0, 500, 256)
    probs[1] = (10, 1024, 30, 200, 10 * 500, 500, -1)
    tiles = []
    for p, (m, n) in enumerate([(10, 384), (30, 200)]):
        for mt in range((m + 127) // 128):
            for nt in range((n + 127) // 128):
                tiles.append((p, mt, nt, 0))
    tiles = torch.tensor(tiles, dtype=torch.int32, device=DEV)
    probs_dev = torch.from_numpy(probs.view(np.uint8).copy()).to(DEV)
    out = torch.full((40, 500), -7.0, device=DEV)
    ops.gemm(a, b, bias=bias, in_dtype=ops.BF16, out=out, out_dtype=ops.F32, problems=probs_dev, tiles=tiles)
    torch.cuda.synchronize()
    ref = torch.full((40, 500), -7.0, device=DEV)
    ref[:10, :384] = (a[:10].double() @ b[256:640].double().t() + bias[256:640].double()).float()
    ref[10:40, :200] = (a[10:40].double() @ b[1024:1224].double().t()).float()
    assert _rel(out, ref) < 2e-5


def test_gemm_simt():
    torch.manual_seed(3)
    n_nodes, ms, ip, ncls = 2, 64, 48, 100
    w = torch.randn(n_nodes, ms * ip, device=DEV)          # [node][a*ip + b]
    wc = torch.randn(ncls, ms, device=DEV)
    bias = torch.randn(ncls, device=DEV)
    out = torch.empty(n_nodes, ncls, ip, device=DEV)
    # out[node][cls][b] = bias[cls] + sum_a wc[cls][a] * relu(w[node][a*ip+b])  (nn.py:757-758)
    ops.gemm_simt(w, 1, ip, wc, ms, 1, bias, out, 1, ip, m=ip, n=ncls, k=ms, relu_a=True, batch=n_nodes,
                  a_bs=ms * ip, d_bs=ncls * ip)
    torch.cuda.synchronize()
    x = torch.relu(w.view(n_nodes, ms, ip)).permute(0, 2, 1)          # n, ip, ms
    ref = (x @ wc.t() + bias).permute(0, 2, 1)
    assert _rel(out, ref) < 1e-5


# --------------------------------------------------------------------------------------------------- graph kernels
GRAPH_SETS = [['resnet18'], ['resnet50', 'vit_b_16', 'swin_v2_t', 'alexnet'], ['efficientnet_v2_l', 'densenet201']]


def _pack(archs):
    recs = [H.graph_records()[a] for a in archs]
    pack = ops.GraphPack([r['n'] for r in recs], edges=[np.asarray(r['edges'], dtype=np.int32) for r in recs],
                         cutoff=50, device=DEV).build()
    return recs, pack


@pytest.mark.parametrize('archs', GRAPH_SETS)
def test_spd_bit_exact(archs):
    recs, pack = _pack(archs)
    torch.cuda.synchronize()
    off = 0
    for g, rec in enumerate(recs):
        spd = pack.spd_matrix(g).cpu().numpy()
        assert H.spd_crc(spd) == rec['spd_crc'], archs[g]          # networkx result recorded from the reference
        A = torch.from_numpy(spd.astype(np.int64))
        din, dout, d0 = O.structural_indices(A)
        n = rec['n']
        assert torch.equal(pack.deg_in[off:off + n].cpu().long(), din)
        assert torch.equal(pack.deg_out[off:off + n].cpu().long(), dout)
        assert torch.equal(pack.dist0[off:off + n].cpu().long(), d0)
        pair = pack.pair_matrix(g).cpu().numpy().astype(np.int64) & 0xFFFF
        assert np.array_equal(pair, spd.astype(np.int64) * 51 + spd.astype(np.int64).T)
        off += n


def test_spd_from_user_matrix_and_cutoffs():
    rec = H.graph_records()['resnet50']
    g = O.graph_from_record(rec)
    pack = ops.GraphPack([rec['n']], spd=[g['A']], cutoff=50, device=DEV).build()
    torch.cuda.synchronize()
    assert np.array_equal(pack.spd_matrix(0).cpu().numpy(), g['A'].astype(np.uint8))
    for cutoff in (1, 2, 5):
        p2 = ops.GraphPack([rec['n']], edges=[np.asarray(rec['edges'])], cutoff=cutoff, device=DEV).build()
        torch.cuda.synchronize()
        assert np.array_equal(p2.spd_matrix(0).cpu().numpy(), O.spd_matrix(g['adj1'], cutoff).astype(np.uint8))


def test_spd_large_synthetic():
    rng = np.random.default_rng(0)
    n = 2500
    edges = [(i - 1, i) for i in range(1, n)]
    for i in range(10, n):
        if rng.random() < 0.3:
            edges.append((int(rng.integers(i - 8, i - 1)), i))
    edges.append((n - 1, 5))          # a cycle
    e = np.asarray(edges, dtype=np.int32)
    pack = ops.GraphPack([n], edges=[e], cutoff=50, device=DEV).build()
    adj = np.zeros((n, n), dtype=np.int64)
    adj[e[:, 0], e[:, 1]] = 1
    ref = O.spd_matrix(adj, 50)
    torch.cuda.synchronize()
    assert np.array_equal(pack.spd_matrix(0).cpu().numpy(), ref.astype(np.uint8))


def _tables(sd):
    return {'embed_op': sd['embed.weight'].to(DEV), 'embed_ch': sd['shape_enc.embed_channel.weight'].to(DEV),
            'embed_sp': sd['shape_enc.embed_spatial.weight'].to(DEV),
            'cent_in': sd['gnn.0.centrality_embed_in.weight'].to(DEV),
            'cent_out': sd['gnn.0.centrality_embed_out.weight'].to(DEV),
            'dist_embed': sd['gnn.0.input_dist_embed.weight'].to(DEV)}


@pytest.mark.parametrize('cfg_name', ['ghn3tiny', 'ghn3xlm16'])
def test_node_features_bit_exact(cfg_name):
    cfg = CONFIGS[cfg_name]
    sd = procedural_state_dict(cfg, 0)
    archs = ['resnet50', 'vit_b_16']
    recs, pack = _pack(archs)
    ops_l, sidx_l, ref_l = [], [], []
    for a, rec in zip(archs, recs):
        model = H.build_model(a)
        g = O.graph_from_record(rec)
        si = O.node_shape_indices(g['node_info'], model, cfg, rec['n'])
        x = O.node_features(sd, g['ops'], si)
        ref_l.append(O.add_structural_embeddings(sd, x, g['A']))
        ops_l.append(g['ops'])
        sidx_l.append(si)
    op = torch.from_numpy(np.concatenate(ops_l).astype(np.int32)).to(DEV)
    sidx = torch.from_numpy(np.concatenate(sidx_l).astype(np.int32)).to(DEV)
    x = ops.node_features(op, sidx, pack, _tables(sd), cfg['hid'])
    torch.cuda.synchronize()
    assert torch.equal(x.cpu(), torch.cat(ref_l))


@pytest.mark.parametrize('cfg_name', ['ghn3tiny', 'ghn3tm8', 'ghn3xlm16'])
def test_edge_lut(cfg_name):
    cfg = CONFIGS[cfg_name]
    sd = procedural_state_dict(cfg, 0)
    lut = ops.edge_lut(sd['gnn.0.attn.edge_embed.embed.weight'].to(DEV), sd['gnn.0.attn.proj_e.0.weight'].to(DEV),
                       sd['gnn.0.attn.proj_e.0.bias'].to(DEV), sd['gnn.0.attn.proj_e.2.weight'].to(DEV),
                       sd['gnn.0.attn.proj_e.2.bias'].to(DEV), 50)
    torch.cuda.synchronize()
    ref = O.edge_bias_lut(sd, 50).reshape(51 * 51, -1).t()
    assert _rel(lut, ref) < 1e-5


@pytest.mark.parametrize('hid', [32, 64, 384, 1024])
def test_layernorm(hid):
    torch.manual_seed(hid)
    x = torch.randn(333, hid, device=DEV) * 3 + 1
    g = torch.randn(hid, device=DEV)
    b = torch.randn(hid, device=DEV)
    ref = torch.nn.functional.layer_norm(x.cpu(), (hid,), g.cpu(), b.cpu(), 1e-5)
    out = ops.layernorm(x, g, b, out_dtype=ops.F32)
    out_bf = ops.layernorm(x, g, b, out_dtype=ops.BF16)
    perm = torch.randperm(333, device=DEV).int()
    perm[::7] = -1
    f32 = torch.empty_like(x)
    out_p = ops.layernorm(x, g, b, out_dtype=ops.TF32, dst_row=perm, out_f32=f32)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-6
    assert _rel(out_bf, ref) < 5e-3
    assert _rel(f32, ref) < 2e-6
    keep = perm.cpu() >= 0
    assert _rel(out_p.cpu()[perm.cpu()[keep].long()], ref[keep]) < 6e-4


@pytest.mark.parametrize('cfg_name,dtype', [('ghn3tiny', ops.BF16), ('ghn3tm8', ops.TF32), ('ghn3lm8', ops.BF16),
                                            ('ghn3xlm16', ops.BF16), ('ghn3xlm16', ops.TF32)])
def test_attention(cfg_name, dtype):
    cfg = CONFIGS[cfg_name]
    C_, H_ = cfg['hid'], cfg['heads']
    D = C_ // H_
    archs = ['resnet18', 'swin_v2_t', 'efficientnet_b0']
    recs, pack = _pack(archs)
    torch.manual_seed(5)
    N = pack.total_nodes
    qkv = torch.randn(N, 3 * C_, device=DEV)
    lut = torch.randn(H_, 51 * 51, device=DEV)
    qkv_in = qkv.bfloat16() if dtype == ops.BF16 else ops.convert(qkv, ops.TF32)
    out = ops.attention(qkv_in, pack, lut, C_, H_, dtype=dtype)
    torch.cuda.synchronize()
    qf = qkv_in.float().cpu()
    off = 0
    for g, rec in enumerate(recs):
        n = rec['n']
        A = pack.spd_matrix(g).cpu().long()
        bias = lut.cpu()[:, (A * 51 + A.t()).reshape(-1)].view(H_, n, n)
        q, k, v = qf[off:off + n].view(n, 3, H_, D).permute(1, 2, 0, 3)
        attn = (q @ k.transpose(-2, -1)) * D ** -0.5 + bias
        ref = (attn.softmax(-1) @ v).transpose(0, 1).reshape(n, C_)
        tol = 1e-2 if dtype == ops.BF16 else 1e-3
        assert _rel(out[off:off + n].float().cpu(), ref) < tol, (cfg_name, g)
        off += n


# --------------------------------------------------------------------------------------------------- scatter
def _desc_array(descs):
    arr = (L.ScatterDesc * len(descs))(*descs)
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(DEV)


def _run_scatter(entries):
    """entries: list of dicts with dst tensor + descriptor fields; fills chunk0 and launches."""
    descs, chunk = [], 0
    for e in entries:
        d = L.ScatterDesc(dst=e['dst'].data_ptr(), src=e['src_ptr'], numel=e['dst'].numel(), chunk0=chunk,
                          t1=e.get('t1', 1), t2=e.get('t2', 1), t3=e.get('t3', 1), so=e['so'], si=e.get('si', 1),
                          ld=e['ld'], ca=e['ca'], ra=e.get('ra', 0), kh_src=e.get('kh_src', 1),
                          kw_src=e.get('kw_src', 1), cy=e.get('cy', 0), cx=e.get('cx', 0), scale=e.get('scale', 1.0),
                          mode=e.get('mode', 0))
        L.fill_fastdiv(d)
        chunk += (e['dst'].numel() + L.SCATTER_CHUNK - 1) // L.SCATTER_CHUNK
        descs.append(d)
    dev = _desc_array(descs)
    ops.scatter(dev, len(descs), chunk)
    torch.cuda.synchronize()


@pytest.mark.parametrize('target,src_shape', [((512, 256, 3, 3), (384, 384, 3, 3)), ((64, 3, 7, 7), (64, 4, 7, 7)),
                                              ((100, 130, 1, 1), (128, 128, 1, 1)), ((96, 1, 5, 5), (128, 4, 5, 5)),
                                              ((1000, 2048), (384, 384, 1, 1)), ((32, 16, 1, 3), (32, 16, 1, 3))])
def test_scatter_conv_matches_tile_normalize(target, src_shape):
    torch.manual_seed(sum(target))
    o, i, kh, kw = src_shape
    P = kh * kw
    w = torch.randn(o, i, kh, kw)                                  # oracle layout (o', i', kh', kw')
    src = w.permute(2, 3, 0, 1).reshape(P, o * i).contiguous().to(DEV)   # device layout [(pos)][a*i'+b]
    ref = O.normalize(O.tile_params(w, target), True)
    dst = torch.empty(target, device=DEV)
    t = tuple(target) + (1,) * (4 - len(target))
    beta = 1.0 if (len(target) > 2 and (target[1] == 1 or target[2] < target[3])) else 2.0
    fan_in = int(np.prod(target[1:]))
    cy = kh // 2 - min(t[2], kh) // 2 if len(target) == 4 else kh // 2
    cx = kw // 2 - min(t[3], kw) // 2 if len(target) == 4 else kw // 2
    _run_scatter([dict(dst=dst, src_ptr=src.data_ptr(), t1=t[1], t2=t[2], t3=t[3], so=o, si=i, ld=o * i, ca=i,
                       kh_src=kh, kw_src=kw, cy=cy, cx=cx, scale=float(np.float32((beta / fan_in) ** 0.5)))])
    assert torch.equal(dst.cpu(), ref) or _rel(dst, ref) < 1e-6


def test_scatter_1d_and_posenc():
    torch.manual_seed(9)
    ms = 384
    w1d = torch.randn(3, 2, ms)
    src = w1d.to(DEV)
    tw, tb, tb2 = torch.empty(1000, device=DEV), torch.empty(1000, device=DEV), torch.empty(37, device=DEV)
    # pos-enc: source (1, i'=384, 14, 14) -> target (1, 197, 768) rows 1.. ; row 0 from a "class token" vector
    pw = torch.randn(1, ms, 14, 14)
    psrc = pw[0].permute(1, 2, 0).reshape(196, ms).contiguous().to(DEV)
    tok = (torch.randn(ms) * 0.02).to(DEV)
    pos = torch.empty(1, 197, 768, device=DEV)
    _run_scatter([
        dict(dst=tw, src_ptr=src[1, 0].data_ptr(), so=ms, ld=0, ca=1, mode=1),
        dict(dst=tb, src_ptr=src[1, 1].data_ptr(), so=ms, ld=0, ca=1, mode=2),
        dict(dst=tb2, src_ptr=src[2, 1].data_ptr(), so=ms, ld=0, ca=1, mode=2),
        dict(dst=pos[0, 0], src_ptr=tok.data_ptr(), so=ms, ld=0, ca=1, mode=0, scale=1.0),
        dict(dst=pos[0, 1:], src_ptr=psrc.data_ptr(), t1=768, so=1 << 30, si=ms, ld=ms, ca=0, ra=1, mode=0, scale=1.0),
    ])
    assert _rel(tw, O.normalize(O.tile_params(w1d[1, 0], (1000,)), True)) < 1e-6
    assert _rel(tb, O.normalize(O.tile_params(w1d[1, 1], (1000,)), False)) < 1e-6
    assert _rel(tb2, O.normalize(O.tile_params(w1d[2, 1], (37,)), False)) < 1e-6
    torch.manual_seed(0)
    ref = O.tile_params(pw, (1, 197, 768))
    assert torch.equal(pos[0, 1:].cpu(), ref[0, 1:])
    assert torch.equal(pos[0, 0].cpu(), tok.cpu().repeat(2))


def test_scatter_bilinear():
    torch.manual_seed(10)
    o, i = 8, 4
    w = torch.randn(o, i, 16, 16)
    src = w.permute(2, 3, 0, 1).reshape(256, o * i).contiguous().to(DEV)
    ref = torch.nn.functional.interpolate(w, (32, 32), mode='bilinear')
    dst = torch.empty(o, i, 32, 32, device=DEV)
    _run_scatter([dict(dst=dst, src_ptr=src.data_ptr(), t1=i, t2=32, t3=32, so=o, si=i, ld=o * i, ca=i, kh_src=16,
                       kw_src=16, mode=3, scale=1.0)])
    assert _rel(dst, ref) < 1e-5


@pytest.mark.parametrize('g', [1, 4, 40, 64, 128])
def test_gemm_grouped_rows_3d_box(g):
    """b_group: compact o' x i' column sub-blocks of the decoder conv.2 weight through a 3-D TMA box."""
    torch.manual_seed(g)
    ms, K, o = 384, 256, 96
    a = torch.randn(70, K, device=DEV).bfloat16()
    w = (torch.randn(ms * ms, K, device=DEV) / 16).bfloat16()
    bias = torch.randn(ms * ms, device=DEV)
    n = o * g
    probs = np.zeros(2, dtype=[('a_row0', 'i4'), ('b_row0', 'i4'), ('m', 'i4'), ('n', 'i4'), ('d_off', 'i8'),
                               ('ldd', 'i4'), ('bias_off', 'i4')])
    probs[0] = (0, 0, 50, n, 0, n, 0)
    probs[1] = (50, 0, 20, n, 50 * n, n, 0)
    tile_n = (128 // g) * g
    tiles = []
    for p, m in enumerate([50, 20]):
        for nt in range((n + tile_n - 1) // tile_n):
            tiles.append((p, 0, nt, 0))
    tiles = torch.tensor(tiles, dtype=torch.int32, device=DEV)
    probs_dev = torch.from_numpy(probs.view(np.uint8).copy()).to(DEV)
    out = torch.zeros(70, n, device=DEV)
    ops.gemm(a, w, bias=bias, in_dtype=ops.BF16, out=out, out_dtype=ops.F32, problems=probs_dev, tiles=tiles,
             b_group=g, b_group_stride=ms, block_n=128)
    torch.cuda.synchronize()
    rows = (torch.arange(o, device=DEV)[:, None] * ms + torch.arange(g, device=DEV)[None, :]).reshape(-1)
    ref = (a.double() @ w[rows].double().t() + bias[rows].double()).float()
    assert _rel(out, ref) < 2e-5


def test_gemm_simt_skinny_rows():
    """bias_class head (nn.py:294): 2 rows per class-bias node, K = max_shape, N = classes."""
    torch.manual_seed(4)
    m, k, n = 4, 384, 1000
    a = torch.randn(m, k, device=DEV)
    w = torch.randn(n, k, device=DEV) / 20
    bias = torch.randn(n, device=DEV)
    out = torch.empty(m, n, device=DEV)
    ops.gemm_simt(a, k, 1, w, k, 1, bias, out, n, 1, m=m, n=n, k=k, relu_a=True)
    torch.cuda.synchronize()
    assert _rel(out, torch.relu(a) @ w.t() + bias) < 1e-5


@pytest.mark.parametrize('dtype', [ops.BF16, ops.TF32, 'tf32x3'])
def test_gemm_grouped_persistent_many_tiles(dtype):
    """Enough tiles (>= 2 per SM) to take the persistent kernel: double-buffered TMEM, ring across tile boundaries.
    'tf32x3' feeds raw fp32 operands through the in-kernel hi/lo split (three MMAs per K step)."""
    x3 = dtype == 'tf32x3'
    if x3:
        dtype = ops.TF32
    torch.manual_seed(11)
    K = 512
    n_prob = 40
    rng = np.random.default_rng(0)
    a = torch.randn(3000, K, device=DEV)
    b = torch.randn(4096, K, device=DEV) / K ** 0.5
    bias = torch.randn(4096, device=DEV)
    if x3:
        a_in, b_in = a, b
    else:
        a_in, b_in = (a.bfloat16(), b.bfloat16()) if dtype == ops.BF16 else (ops.convert(a, ops.TF32), ops.convert(b, ops.TF32))
    probs = np.zeros(n_prob, dtype=[('a_row0', 'i4'), ('b_row0', 'i4'), ('m', 'i4'), ('n', 'i4'), ('d_off', 'i8'),
                                    ('ldd', 'i4'), ('bias_off', 'i4')])
    tiles, off, specs = [], 0, []
    for p in range(n_prob):
        m = int(rng.integers(1, 400)); n = int(rng.integers(1, 700))
        a0 = int(rng.integers(0, 3000 - m)); b0 = int(rng.integers(0, 4096 - n))
        ldd = n + int(rng.integers(0, 9))
        probs[p] = (a0, b0, m, n, off, ldd, b0 if p % 2 == 0 else -1)
        specs.append((a0, b0, m, n, off, ldd, p % 2 == 0))
        off += m * ldd
        for mt in range((m + 127) // 128):
            for nt in range((n + 127) // 128):
                tiles.append((p, mt, nt, 0))
    assert len(tiles) >= 2 * 148
    out = torch.full((off,), 3.0, device=DEV)
    ops.gemm(a_in, b_in, bias=bias, act=ops.ACT_RELU, in_dtype=dtype, out=out, out_dtype=ops.F32,
             problems=torch.from_numpy(probs.view(np.uint8).copy()).to(DEV),
             tiles=torch.tensor(tiles, dtype=torch.int32, device=DEV), x3=x3)
    torch.cuda.synchronize()
    for (a0, b0, m, n, o, ldd, hb) in specs:
        ref = a_in[a0:a0 + m].double() @ b_in[b0:b0 + n].double().t()
        if hb:
            ref = ref + bias[b0:b0 + n].double()
        ref = torch.relu(ref).float()
        got = out[o:o + m * ldd].view(m, ldd)
        assert _rel(got[:, :n], ref) < (3e-5 if x3 else 2e-5)
        assert bool((got[:, n:] == 3.0).all())          # padding columns untouched


@pytest.mark.parametrize('dtype', [ops.BF16, ops.TF32])
@pytest.mark.parametrize('g', [0, 4])
def test_gemm_grouped_swap_ab(dtype, g):
    """swap_ab: weights on the 128 UMMA-M rows, 64 activation rows per tile, row-mapped column-coalesced stores."""
    torch.manual_seed(21 + g)
    K, ms = 384, 96
    rng = np.random.default_rng(1)
    a = torch.randn(500, K, device=DEV)
    w = torch.randn(ms * ms, K, device=DEV) / K ** 0.5
    bias = torch.randn(ms * ms, device=DEV)
    a_in, w_in = (a.bfloat16(), w.bfloat16()) if dtype == ops.BF16 else (ops.convert(a, ops.TF32), ops.convert(w, ops.TF32))
    n_prob = 30
    probs = np.zeros(n_prob, dtype=[('a_row0', 'i4'), ('b_row0', 'i4'), ('m', 'i4'), ('n', 'i4'), ('d_off', 'i8'),
                                    ('ldd', 'i4'), ('bias_off', 'i4')])
    tiles, rowmap, specs = [], [], []
    total_rows = 0
    ldd = 1536
    for p in range(n_prob):
        m = int(rng.integers(1, 150))
        a0 = int(rng.integers(0, 500 - m))
        if g == 0:
            n = int(rng.integers(1, 1500)); b0 = int(rng.integers(0, ms * ms - n)); tile_n = 128
            rows_w = np.arange(b0, b0 + n)
        else:
            o = int(rng.integers(1, ms)); n = o * g; b0 = 0; tile_n = (128 // g) * g
            rows_w = (np.arange(o)[:, None] * ms + np.arange(g)[None, :]).reshape(-1)
        probs[p] = (a0, b0, m, n, len(rowmap), ldd, b0)
        perm = total_rows + rng.permutation(m)
        rowmap.extend(perm.tolist())
        specs.append((a0, m, rows_w, perm))
        total_rows += m
        for mt in range((m + 63) // 64):
            for nt in range((n + tile_n - 1) // tile_n):
                tiles.append((p, mt, nt, 0))
    out = torch.full((total_rows, ldd), -5.0, device=DEV, dtype=torch.bfloat16 if dtype == ops.BF16 else torch.float32)
    ops.gemm(a_in, w_in, bias=bias, act=ops.ACT_RELU, in_dtype=dtype, out=out, out_dtype=dtype if dtype == ops.BF16 else ops.F32,
             problems=torch.from_numpy(probs.view(np.uint8).copy()).to(DEV),
             tiles=torch.tensor(tiles, dtype=torch.int32, device=DEV),
             rowmap=torch.tensor(rowmap, dtype=torch.int32, device=DEV), swap_ab=True, b_group=g,
             b_group_stride=ms if g else 0, b_dynamic=False)
    torch.cuda.synchronize()
    for (a0, m, rows_w, perm) in specs:
        rw = torch.from_numpy(rows_w).to(DEV)
        ref = torch.relu(a_in[a0:a0 + m].double() @ w_in[rw].double().t() + bias[rw].double()).float()
        got = out[torch.from_numpy(perm).to(DEV)].float()
        assert _rel(got[:, :len(rows_w)], ref) < (6e-3 if dtype == ops.BF16 else 2e-5)
        assert bool((got[:, len(rows_w):] == -5.0).all())


@pytest.mark.parametrize('dtype,x3', [(ops.BF16, False), (ops.TF32, True)])
@pytest.mark.parametrize('splits', [0, 1, 4])
def test_gemm_fused_layernorm(dtype, x3, splits):
    """Residual GEMM + LayerNorm of the finished rows by the last-arriving CTA (graphormer.py:239-241)."""
    import ctypes as C_
    torch.manual_seed(31 + splits)
    m, n, k = 457, 384, 1536
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    bias = torch.randn(n, device=DEV)
    g, be = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    a_in, b_in = (a, b) if x3 else (a.bfloat16(), b.bfloat16())
    x = torch.randn(m, n, device=DEV)
    x_ref = (x.double() + a_in.double() @ b_in.double().t() + bias.double()).float()
    ln_ref = torch.nn.functional.layer_norm(x_ref, (n,), g, be, 1e-5)
    ln_out = torch.full((m, n), 7.0, device=DEV, dtype=torch.float32 if x3 else torch.bfloat16)
    counters = torch.zeros(8, dtype=torch.int32, device=DEV)
    for rep in range(2):                       # the counters must come back to zero
        xx = x.clone()
        ga = L.GemmArgs(a=L.ptr(a_in), a_rows=m, lda=k, b=L.ptr(b_in), b_rows=n, ldb=k, k=k, in_dtype=dtype,
                        d=L.ptr(xx), out_dtype=ops.F32, bias=L.ptr(bias), act=0, accumulate=1, k_splits=splits,
                        tf32_x3=int(x3), b_dynamic=1, ln_out=L.ptr(ln_out), ln_gamma=L.ptr(g), ln_beta=L.ptr(be),
                        ln_counters=L.ptr(counters), ln_out_dtype=ops.F32 if x3 else ops.BF16)
        ga.single = L.GemmProblem(a_row0=0, b_row0=0, m=m, n=n, d_off=0, ldd=n, bias_off=0)
        L.call('gemm', ga, L.current_stream())
        torch.cuda.synchronize()
        assert _rel(xx, x_ref) < 2e-5
        assert _rel(ln_out, ln_ref) < (1e-2 if not x3 else 2e-5)
        assert int(counters.abs().sum()) == 0


@pytest.mark.parametrize('cfg_name,archs', [('ghn3xlm16', ['efficientnet_v2_l', 'resnet18']), ('ghn3lm8', ['swin_v2_t']),
                                            ('ghn3tm8', ['resnet50', 'efficientnet_b0', 'alexnet']),
                                            ('ghn3sm8', ['densenet201'])])
def test_attention_tcgen05(cfg_name, archs):
    """The tcgen05 attention kernel (S and P.V in TMEM; large graphs) against the fp32 reference and against the
    mma.sync kernel, with the softmax statistics it keeps for the backward pass."""
    cfg = CONFIGS[cfg_name]
    C_, H_ = cfg['hid'], cfg['heads']
    D = C_ // H_
    recs, pack = _pack(archs)
    torch.manual_seed(7)
    N = pack.total_nodes
    qkv = torch.randn(N, 3 * C_, device=DEV).bfloat16()
    lut = torch.randn(H_, 51 * 51, device=DEV)
    lib = L.load()
    old = lib.ghn3_set_attention_tc_min(1 << 30)
    try:
        lse_a = torch.zeros(H_, N, device=DEV)
        ref_kernel = ops.attention(qkv, pack, lut, C_, H_, dtype=ops.BF16, lse2=lse_a)
        lib.ghn3_set_attention_tc_min(0)
        lse_b = torch.zeros(H_, N, device=DEV)
        out = ops.attention(qkv, pack, lut, C_, H_, dtype=ops.BF16, lse2=lse_b)
        torch.cuda.synchronize()
    finally:
        lib.ghn3_set_attention_tc_min(old)
    assert _rel(out, ref_kernel) < 1e-2
    assert _rel(lse_b, lse_a) < 1e-3
    qf = qkv.float().cpu()
    off = 0
    for g, rec in enumerate(recs):
        n = rec['n']
        A = pack.spd_matrix(g).cpu().long()
        bias = lut.cpu()[:, (A * 51 + A.t()).reshape(-1)].view(H_, n, n)
        q, k, v = qf[off:off + n].view(n, 3, H_, D).permute(1, 2, 0, 3)
        attn = (q @ k.transpose(-2, -1)) * D ** -0.5 + bias
        ref = (attn.softmax(-1) @ v).transpose(0, 1).reshape(n, C_)
        assert _rel(out[off:off + n].float().cpu(), ref) < 1e-2, (cfg_name, g)
        off += n
