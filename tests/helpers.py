"""Shared helpers for the tests: golden fixtures, model construction, error metrics."""
import gzip
import json
import os
import zlib

import numpy as np
import torch
import torchvision.models as tvm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')

_graphs = None


def graph_records():
    global _graphs
    if _graphs is None:
        with gzip.open(os.path.join(GOLDEN, 'graphs_tv.json.gz'), 'rt') as f:
            _graphs = json.load(f)
        _graphs.update(cellnet_records()['graphs'])          # 'cellnet0' .. 'cellnet7'
    return _graphs


def cellnet_records():
    """tests/golden/graphs_cellnets.json.gz: the reference's tracer on NetGenerator(seed) networks."""
    with gzip.open(os.path.join(GOLDEN, 'graphs_cellnets.json.gz'), 'rt') as f:
        return json.load(f)


def build_model(name):
    """Same construction as tests/golden/make_golden.py:build_model (seeded default init)."""
    if name.startswith('cellnet'):
        # the i-th network of ghn3_b200.deepnets.NetGenerator(seed of the fixture); default (unseeded) weight init is
        # irrelevant: every parameter is predicted
        from ghn3_b200.deepnets import NetGenerator
        gen = NetGenerator(seed=cellnet_records()['seed'])
        for _ in range(int(name[7:]) + 1):
            net = gen.sample_net()
        net.expected_input_sz = 64
        return net
    kw = {'init_weights': False} if name in ['googlenet', 'inception_v3'] else {}
    torch.manual_seed(0)
    m = getattr(tvm, name)(**kw)
    if name == 'inception_v3':
        m.expected_input_sz = 299
    return m


def load_pred(cfg, arch):
    with open(os.path.join(GOLDEN, 'pred_%s_%s.json' % (cfg, arch))) as f:
        return json.load(f)


def load_emb(cfg, arch):
    return torch.from_numpy(np.load(os.path.join(GOLDEN, 'emb_%s_%s.npy' % (cfg, arch))))


def fingerprint(t):
    t = t.detach().double().reshape(-1).cpu()
    n = t.numel()
    idx = np.unique(np.linspace(0, n - 1, 16).astype(np.int64))
    return {'numel': n, 'sum': float(t.sum()), 'abs': float(t.abs().sum()), 'sq': float((t * t).sum()),
            'idx': [int(i) for i in idx], 'val': [float(t[i]) for i in idx]}


def spd_crc(A):
    return zlib.crc32(np.asarray(A).astype(np.uint8).tobytes())


def max_rel_err(a, b):
    """max |a - b| / max |b| : the 'max relative error' of BASELINE.json's north_star, per tensor."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / max(denom, 1e-30)


def elem_rel_err(a, b, floor=1e-2):
    """Element-wise relative error max_i |a_i - b_i| / max(|b_i|, floor * max|b|): the stricter reading of "max relative
    error" (VERDICT r1). Elements far below the tensor's scale are measured against `floor` x that scale -- an
    unfloored ratio is meaningless for values that are themselves rounding residue."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(b.abs().max().item(), 1e-30)
    return ((a - b).abs() / b.abs().clamp_min(floor * scale)).max().item()
