"""
Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference sources
(/root/reference/ghn3) in this container. The reference's un-vendored `ppuda` dependency is provided by the
restatement in tests/golden/ref_shim; h5py is stubbed.

This script can only run where /root/reference is mounted (the build container). Tests and the GPU box consume
only the committed outputs:

  graphs_tv.json.gz        graph structure of every torchvision classification constructor:
                           op ids, 1-hop edges, node_info, crc32 of the networkx SPD matrix (uint8, row-major)
  pred_<cfg>_<arch>.json   per-tensor fingerprints of the parameters the reference predicts with
                           ghn3_b200.weights.procedural_state_dict(cfg, seed=0) loaded into the reference GHN3
  emb_<cfg>_<arch>.npy     node embeddings after the Graphormer stack + final LayerNorm (return_embeddings=True)

  graphs_cellnets.json.gz  the same for the first 8 networks of ghn3_b200.deepnets.NetGenerator(seed=0)

Usage:  python tests/golden/make_golden.py graphs | cellnets | msa | preds | all
"""
import gzip
import inspect
import json
import os
import sys
import time
import warnings
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference', REPO]
warnings.filterwarnings('ignore')

import numpy as np
import torch
import torchvision.models as models

import ghn3 as ref                                   # the reference package
from ghn3_b200.weights import CONFIGS, procedural_state_dict

PRED_SEED = 1234                                     # torch seed set right before every reference forward


def tv_model_names():
    """Same enumeration as the reference's eval_ghn.py:76-90."""
    names = []
    for m in dir(models):
        if m[0].isupper() or m.startswith('_') or m.startswith('get') or m == 'list_models' or \
                not inspect.isfunction(getattr(models, m)):
            continue
        names.append(m)
    return names


def build_model(name):
    if name.startswith('cellnet'):                    # i-th network of ghn3_b200.deepnets.NetGenerator(seed=0)
        from tests import helpers
        return helpers.build_model(name)
    kw = {'init_weights': False} if name in ['googlenet', 'inception_v3'] else {}
    torch.manual_seed(0)
    m = getattr(models, name)(**kw)
    if name == 'inception_v3':
        m.expected_input_sz = 299
    return m


def ser_sz(sz):
    return None if sz is None else [int(v) for v in sz]


def graph_record(name, model=None):
    model = build_model(name) if model is None else model
    t0 = time.time()
    g = ref.Graph(model, ve_cutoff=50, verbose=False)
    dt = time.time() - t0
    A = g._Adj.numpy()
    assert A.max() <= 50 and A.min() >= 0
    e = np.argwhere(A == 1)
    spd = A.astype(np.uint8)
    rec = {
        'n': int(g.n_nodes),
        'ops': [int(v) for v in g.node_feat[:, 0]],
        'edges': [[int(a), int(b)] for a, b in e],
        'node_info': [[[int(r[0]), r[1], r[2], ser_sz(r[3]), bool(r[4]), bool(r[5])] for r in cell]
                      for cell in g.node_info],
        'spd_crc': zlib.crc32(spd.tobytes()),
        'spd_nnz': int((A > 0).sum()),
        'spd_max': int(A.max()),
        'n_params': int(sum(p.numel() for p in model.parameters())),
        'trace_sec': round(dt, 3),
    }
    return rec


def make_graphs():
    out = {}
    for i, name in enumerate(tv_model_names()):
        rec = graph_record(name)
        out[name] = rec
        print('%3d %-22s N=%4d edges=%4d nnz=%6d max=%2d trace=%.2fs' % (
            i, name, rec['n'], len(rec['edges']), rec['spd_nnz'], rec['spd_max'], rec['trace_sec']), flush=True)
    with gzip.open(os.path.join(HERE, 'graphs_tv.json.gz'), 'wt') as f:
        json.dump(out, f, separators=(',', ':'))
    print('wrote %d graphs' % len(out))


def make_cellnet_graphs(n=8, seed=0):
    """graphs_cellnets.json.gz: the reference's tracer on the first `n` networks of
    ghn3_b200.deepnets.NetGenerator(seed) (DeepNets-1M-style cell networks with `_n_cells`, per-cell node_info)."""
    from ghn3_b200.deepnets import NetGenerator
    gen = NetGenerator(seed=seed)
    out = {}
    for i in range(n):
        net = gen.sample_net()
        net.expected_input_sz = 64
        rec = graph_record('cellnet%d' % i, model=net)
        out['cellnet%d' % i] = rec
        print('%3d cellnet N=%4d edges=%4d nnz=%6d max=%2d trace=%.2fs' % (
            i, rec['n'], len(rec['edges']), rec['spd_nnz'], rec['spd_max'], rec['trace_sec']), flush=True)
    with gzip.open(os.path.join(HERE, 'graphs_cellnets.json.gz'), 'wt') as f:
        json.dump({'seed': seed, 'graphs': out}, f, separators=(',', ':'))
    print('wrote %d cell-network graphs' % len(out))


def make_msa_graphs(n=3, seed=5):
    """graphs_cellnets_msa.json.gz: the reference's tracer on the first `n` networks of
    NetGenerator(seed, with_msa=True) that contain the 'msa' primitive (ghn3/ops.py:302)."""
    from ghn3_b200.deepnets import NetGenerator
    gen = NetGenerator(seed=seed, with_msa=True, max_params=8e6)
    out, picked, i = {}, [], -1
    while len(out) < n:
        net = gen.sample_net()
        i += 1
        g = net.net_args['genotype']
        if not any(e[0] == 'msa' for e in g['normal'] + g['reduce']):
            continue
        net.expected_input_sz = 64
        rec = graph_record('cellnet_msa%d' % len(out), model=net)
        out['cellnet_msa%d' % len(out)] = rec
        picked.append(i)
        print('%3d msa cellnet N=%4d edges=%4d nnz=%6d max=%2d trace=%.2fs' % (
            i, rec['n'], len(rec['edges']), rec['spd_nnz'], rec['spd_max'], rec['trace_sec']), flush=True)
    with gzip.open(os.path.join(HERE, 'graphs_cellnets_msa.json.gz'), 'wt') as f:
        json.dump({'seed': seed, 'stream_index': picked, 'graphs': out}, f, separators=(',', ':'))
    print('wrote %d msa cell-network graphs' % len(out))


def fingerprint(t):
    t = t.detach().double().reshape(-1)
    n = t.numel()
    idx = np.unique(np.linspace(0, n - 1, 16).astype(np.int64))
    return {'numel': n, 'sum': float(t.sum()), 'abs': float(t.abs().sum()), 'sq': float((t * t).sum()),
            'idx': [int(i) for i in idx], 'val': [float(t[i]) for i in idx]}


def make_pred(cfg_name, arch, save_emb=True):
    cfg = CONFIGS[cfg_name]
    ghn = ref.GHN3(max_shape=cfg['max_shape'], num_classes=cfg['num_classes'], hid=cfg['hid'], heads=cfg['heads'],
                   layers=cfg['layers'], weight_norm=True, ve=True, layernorm=cfg['layernorm']).eval()
    ghn.load_state_dict(procedural_state_dict(cfg, seed=0))
    model = build_model(arch)
    graph = ref.Graph(model, ve_cutoff=50, verbose=False)
    torch.manual_seed(PRED_SEED)
    t0 = time.time()
    with torch.no_grad():
        model, emb = ghn(model, graph, return_embeddings=True, bn_track_running_stats=True, keep_grads=False)
    dt = time.time() - t0
    rec = {'cfg': cfg_name, 'arch': arch, 'seed': 0, 'torch_seed': PRED_SEED, 'forward_sec': round(dt, 3),
           'n_nodes': int(graph.n_nodes), 'tensors': {}}
    for n, p in model.named_parameters():
        fp = fingerprint(p)
        fp['shape'] = [int(v) for v in p.shape]
        rec['tensors'][n] = fp
    with open(os.path.join(HERE, 'pred_%s_%s.json' % (cfg_name, arch)), 'w') as f:
        json.dump(rec, f)
    if save_emb:
        np.save(os.path.join(HERE, 'emb_%s_%s.npy' % (cfg_name, arch)), emb.numpy().astype(np.float32))
    print('pred %s %s: %d tensors, %.2fs' % (cfg_name, arch, len(rec['tensors']), dt), flush=True)


PRED_CASES = [('ghn3tiny', 'resnet18'), ('ghn3tiny', 'squeezenet1_1'), ('ghn3tiny', 'mobilenet_v3_small'),
              ('ghn3tiny', 'vit_b_32'), ('ghn3tiny', 'swin_v2_t'), ('ghn3tiny', 'convnext_tiny'),
              ('ghn3tiny', 'cellnet7'), ('ghn3tm8', 'resnet50'),
              ('ghn3xlm16', 'vit_b_16'), ('ghn3xlm16', 'convnext_base')]


def make_preds():
    for cfg_name, arch in PRED_CASES:
        make_pred(cfg_name, arch)


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('graphs', 'all'):
        make_graphs()
    if what in ('cellnets', 'all'):
        make_cellnet_graphs()
    if what in ('msa', 'all'):
        make_msa_graphs()
    if what in ('preds', 'all'):
        make_preds()
    if what.startswith('pred:'):                      # one case, e.g. pred:ghn3tiny:cellnet7
        _, cfg_name, arch = what.split(':')
        make_pred(cfg_name, arch)
