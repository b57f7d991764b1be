"""The persistent fused Graphormer-stack kernel (ghn3_graphormer_fused) against a PyTorch fp32 reference with the
same bf16 rounding points, stage by stage (stop_after) and end to end, and against the unfused stack."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ghn3_b200 import _lib as L
from ghn3_b200 import ops
from tests import helpers as H

DEV = 'cuda'


def _pack(archs):
    recs = [H.graph_records()[a] for a in archs]
    pack = ops.GraphPack([r['n'] for r in recs], edges=[np.asarray(r['edges'], dtype=np.int32) for r in recs],
                         cutoff=50, device=DEV).build()
    return recs, pack


def _weights(C, layers, seed):
    g = torch.Generator(device='cpu').manual_seed(seed)
    R = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)
    per = []
    for _ in range(layers):
        per.append(dict(ln1_w=1 + R(C, scale=0.1), ln1_b=R(C, scale=0.1), w_qkv=R(3 * C, C, scale=C ** -0.5),
                        w_out=R(C, C, scale=C ** -0.5), b_out=R(C, scale=0.1), ln2_w=1 + R(C, scale=0.1),
                        ln2_b=R(C, scale=0.1), w_ff1=R(4 * C, C, scale=C ** -0.5), b_ff1=R(4 * C, scale=0.1),
                        w_ff2=R(C, 4 * C, scale=(4 * C) ** -0.5), b_ff2=R(C, scale=0.1)))
    stack = {k: torch.cat([p[k] for p in per]).bfloat16().contiguous() for k in ('w_qkv', 'w_out', 'w_ff1', 'w_ff2')}
    return per, stack


def _bf(t):
    return t.bfloat16().float()


def _ref_layer(x, p, stackl, pack, recs, lut, C, Hh):
    """Returns the intermediates (qkv, ao, x_after_proj, ff, x_out) of one layer, fp32 math, bf16 rounding points."""
    D = C // Hh
    F = torch.nn.functional
    h = _bf(F.layer_norm(x, (C,), p['ln1_w'], p['ln1_b'], 1e-5))
    qkv = _bf(h @ stackl['w_qkv'].float().t())
    ao = torch.empty(x.shape[0], C, device=DEV)
    off = 0
    for g, rec in enumerate(recs):
        n = rec['n']
        A = pack.spd_matrix(g).long()
        bias = lut[:, (A * 51 + A.t()).reshape(-1)].view(Hh, n, n)
        q, k, v = qkv[off:off + n].view(n, 3, Hh, D).permute(1, 2, 0, 3)
        att = ((q @ k.transpose(-2, -1)) * D ** -0.5 + bias).softmax(-1)
        ao[off:off + n] = (att @ v).transpose(0, 1).reshape(n, C)
        off += n
    ao = _bf(ao)
    x1 = x + ao @ stackl['w_out'].float().t() + p['b_out']
    h2 = _bf(F.layer_norm(x1, (C,), p['ln2_w'], p['ln2_b'], 1e-5))
    ff = _bf(F.gelu(h2 @ stackl['w_ff1'].float().t() + p['b_ff1']))
    x2 = x1 + ff @ stackl['w_ff2'].float().t() + p['b_ff2']
    return qkv, ao, x1, ff, x2


def _setup(C, Hh, layers, archs, seed=0):
    recs, pack = _pack(archs)
    per, stack = _weights(C, layers, seed)
    tab = ops.layer_table([{k: v for k, v in p.items() if not k.startswith('w_')} for p in per], DEV)
    torch.manual_seed(seed + 1)
    N = pack.total_nodes
    x0 = torch.randn(N, C, device=DEV)
    lut = torch.randn(Hh, 51 * 51, device=DEV)
    fg = ops.FusedGraphormer(C, Hh, layers, stack, tab, N, DEV)
    fg.bind(pack, lut)
    fg._tab_src = per
    return recs, pack, per, stack, x0, lut, fg


def _slice(stack, l, C):
    return {'w_qkv': stack['w_qkv'][l * 3 * C:(l + 1) * 3 * C], 'w_out': stack['w_out'][l * C:(l + 1) * C],
            'w_ff1': stack['w_ff1'][l * 4 * C:(l + 1) * 4 * C], 'w_ff2': stack['w_ff2'][l * C:(l + 1) * C]}


def _rel(a, b):
    return H.max_rel_err(a.float(), b.float())


@pytest.mark.parametrize('C,Hh,archs', [(384, 16, ['vit_b_16', 'convnext_base']), (64, 8, ['resnet18']),
                                       (256, 16, ['resnet50', 'alexnet', 'swin_v2_t'])])
def test_fused_stage_by_stage(C, Hh, archs):
    """Layer 0 and 1, one stage at a time: pins which stage is wrong if anything is."""
    recs, pack, per, stack, x0, lut, fg = _setup(C, Hh, 2, archs)
    refs = []
    x = x0
    for l in range(2):
        out = _ref_layer(x, per[l], _slice(stack, l, C), pack, recs, lut, C, Hh)
        refs.append(out)
        x = out[4]
    for stop in range(1, 11):
        fg.x.copy_(x0)
        fg.run(stop_after=stop)
        torch.cuda.synchronize()
        l, s = (stop - 1) // 5, (stop - 1) % 5
        got = [fg.qkv[l & 1], fg.ao, fg.x, fg.ff, fg.x][s]
        err = _rel(got, refs[l][s])
        assert err < 2e-2, ('stage', stop, 'layer', l, ['qkv', 'attn', 'proj', 'ff1', 'ff2'][s], err)


@pytest.mark.parametrize('C,Hh,layers,archs', [(384, 16, 24, ['vit_b_16', 'convnext_base']),
                                              (384, 16, 24, ['efficientnet_v2_l']),
                                              (256, 16, 12, ['resnet50']),
                                              (128, 16, 5, ['resnet18', 'alexnet']),
                                              (64, 8, 3, ['resnet50', 'vit_b_16', 'swin_v2_t', 'alexnet'])])
def test_fused_full_stack(C, Hh, layers, archs):
    recs, pack, per, stack, x0, lut, fg = _setup(C, Hh, layers, archs, seed=3)
    x = x0
    for l in range(layers):
        x = _ref_layer(x, per[l], _slice(stack, l, C), pack, recs, lut, C, Hh)[4]
    for rep in range(3):                       # re-entrancy: counters are reset by every call
        fg.x.copy_(x0)
        out = fg.run()
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        err = float((out - x).norm() / x.norm())
        assert err < 1e-2, (C, layers, archs, rep, err)
        assert _rel(out, x) < 3e-2


def test_fused_few_ctas_multi_round():
    """Fewer CTAs than tiles: every CTA walks several tiles per stage (the general persistent schedule)."""
    C, Hh, layers = 384, 16, 4
    recs, pack, per, stack, x0, lut, fg = _setup(C, Hh, layers, ['efficientnet_b0', 'resnet50'], seed=5)
    x = x0
    for l in range(layers):
        x = _ref_layer(x, per[l], _slice(stack, l, C), pack, recs, lut, C, Hh)[4]
    for ctas in (148, 40, 7):
        fg.x.copy_(x0)
        out = fg.run(max_ctas=ctas)
        torch.cuda.synchronize()
        err = float((out - x).norm() / x.norm())
        assert err < 1e-2, (ctas, err)
