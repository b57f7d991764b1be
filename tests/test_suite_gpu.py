"""BASELINE.json config 3/4 coverage: every torchvision classification architecture through the CUDA path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ghn3_b200 import GHN3, Graph
from ghn3_b200.plan import ModelPlan
from ghn3_b200.weights import CONFIGS, procedural_state_dict
from oracle import ghn3_oracle as O
from tests import helpers as H

DEV = 'cuda'
ORACLE_CHECKED = ['efficientnet_v2_l', 'regnet_y_400mf', 'densenet201', 'inception_v3', 'maxvit_t', 'swin_v2_s',
                  'mnasnet1_0', 'shufflenet_v2_x1_0', 'vgg16_bn', 'googlenet', 'wide_resnet50_2', 'vit_l_32']


@pytest.mark.parametrize('dtype,tol,checked', [('bf16', 2e-2, ORACLE_CHECKED), ('tf32', 1e-3, ORACLE_CHECKED[:6])])
def test_all_torchvision_models_lm8(dtype, tol, checked):
    """ghn3lm8-sized GHN (random-init) predicts all 80 architectures; a subset is compared tensor by tensor with the
    oracle (<= 2e-2 bf16, <= 1e-3 tf32), every model must have all its matched parameters written with finite values."""
    cfg = CONFIGS['ghn3lm8']
    sd = procedural_state_dict(cfg, 0)
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=dtype)
    ghn.load_state_dict(sd)
    ghn = ghn.to(DEV).eval()
    recs = H.graph_records()
    failures = []
    for arch, rec in recs.items():
        model = H.build_model(arch).to(DEV)
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        with torch.no_grad():
            ghn(model, Graph.from_record(rec))
        torch.cuda.synchronize()
        plan = ModelPlan(Graph.from_record(rec), model, cfg)
        predicted = {id(getattr(m, a)) for (m, a, _, _) in
                     __import__('ghn3_b200.plan', fromlist=['BatchPlan']).BatchPlan([plan], cfg).desc_targets}
        for n, p in model.named_parameters():
            if not torch.isfinite(p).all():
                failures.append((arch, n, 'non-finite'))
            if id(p) in predicted and torch.equal(p, before[n]):
                failures.append((arch, n, 'not written'))
        if arch in checked:
            ref = H.build_model(arch)
            O.predict(sd, cfg, ref, O.graph_from_record(rec))
            refp = dict(ref.named_parameters())
            for n, p in model.named_parameters():
                r = refp[n]
                if n.endswith('pos_embedding'):
                    p, r = p[:, 1:], r[:, 1:]
                err = H.max_rel_err(p, r)
                if err > tol:
                    failures.append((arch, n, err))
        del model
    assert not failures, failures[:20]


@pytest.mark.parametrize('cfg_name,n,dtype,tol', [('ghn3tm8', 1024, 'bf16', 2e-2), ('ghn3tm8', 3000, 'bf16', 2e-2),
                                                  ('ghn3tm8', 3000, 'tf32', 1e-3), ('ghn3xlm16', 1024, 'bf16', 2e-2),
                                                  ('ghn3xlm16', 1024, 'tf32', 1e-3)])
def test_large_synthetic_graph_stack_runner(cfg_name, n, dtype, tol):
    """Config 4 as SURVEY 8d specifies it: synthetic DAGs of >= 1024 nodes, also at ghn3xlm16 and in the accurate mode;
    ghn3_b200.synthetic.StackRunner (node features + SPD + stack + final LayerNorm) vs the oracle."""
    from ghn3_b200.synthetic import StackRunner, synthetic_dag
    cfg = CONFIGS[cfg_name]
    sd = procedural_state_dict(cfg, 0)
    edges, op = synthetic_dag(n, seed=0)
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=dtype)
    ghn.load_state_dict(sd)
    ghn = ghn.to(DEV).eval()
    sr = StackRunner(ghn, [n], [edges], [op], DEV)
    emb = sr.run()
    torch.cuda.synchronize()
    adj = np.zeros((n, n), dtype=np.int64)
    adj[edges[:, 0], edges[:, 1]] = 1
    A = O.spd_matrix(adj, 50)
    assert np.array_equal(sr.pack.spd_matrix(0).cpu().numpy(), A.astype(np.uint8))
    tabs = sr.sidx.cpu().numpy().astype(np.int64)
    x0 = O.node_features(sd, op, tabs)
    ref = O.graphormer_stack(sd, cfg, x0, A)
    assert H.max_rel_err(emb, ref) < tol, H.max_rel_err(emb, ref)


@pytest.mark.parametrize('n', [1024, 3000])
def test_large_synthetic_graph_stack(n):
    """Config 4: thousands of nodes. Synthetic DAG (chain + random skips), Graphormer stack vs the oracle."""
    from ghn3_b200 import ops, _lib as L
    cfg = CONFIGS['ghn3tm8']
    sd = procedural_state_dict(cfg, 0)
    rng = np.random.default_rng(0)
    edges = [(i - 1, i) for i in range(1, n)]
    for i in range(10, n):
        if rng.random() < 0.3:
            edges.append((int(rng.integers(i - 8, i - 1)), i))
    e = np.asarray(edges, dtype=np.int32)
    op = rng.integers(0, 15, size=n).astype(np.int32)
    pack = ops.GraphPack([n], edges=[e], cutoff=50, device=DEV, op=op).build()
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
    ghn.load_state_dict(sd)
    ghn = ghn.to(DEV).eval()
    w = ghn._device_weights()
    C_ = cfg['hid']
    sidx = torch.from_numpy(np.tile(np.array([391, 391, 10, 10], dtype=np.int32), (n, 1))).to(DEV)
    x = ops.node_features(pack.op_dev, sidx, pack, w['tables'], C_)
    emb = torch.empty(n, C_, device=DEV)
    tdt = torch.bfloat16
    h, qkv, ff = (torch.empty(n, k * C_, dtype=tdt, device=DEV) for k in (1, 3, 4))
    dec = torch.empty(n, C_, dtype=tdt, device=DEV)
    lut = ghn._lut(w, 50)
    ga = L.GraphormerArgs(hid=C_, heads=cfg['heads'], layers=cfg['layers'], dtype=ops.BF16, layers_host=w['layers'],
                          ln_w=L.ptr(w['ln_w']), ln_b=L.ptr(w['ln_b']), n_graphs=1, total_nodes=n, max_nodes=n,
                          lut_size=lut.shape[1], node_off=L.ptr(pack.d['node_off']), mat_off=L.ptr(pack.d['mat_off']),
                          pair=L.ptr(pack.pair), lut=L.ptr(lut), x=L.ptr(x), h=L.ptr(h), qkv=L.ptr(qkv), ff=L.ptr(ff),
                          dec_in=L.ptr(dec), dec_dtype=ops.BF16, emb_f32=L.ptr(emb))
    L.call('graphormer_stack', ga, L.current_stream())
    torch.cuda.synchronize()
    adj = np.zeros((n, n), dtype=np.int64)
    adj[e[:, 0], e[:, 1]] = 1
    A = O.spd_matrix(adj, 50)
    x0 = O.node_features(sd, op, np.tile(np.array([391, 391, 10, 10]), (n, 1)))
    ref = O.graphormer_stack(sd, cfg, x0, A)
    assert H.max_rel_err(emb, ref) < 2e-2
