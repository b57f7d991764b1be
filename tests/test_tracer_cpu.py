"""Our host tracer against graphs produced by the reference's tracer (tests/golden/graphs_tv.json.gz): bit-exact."""
import numpy as np
import pytest
import torch

from ghn3_b200.graph import Graph
from tests import helpers as H

FAST = ['alexnet', 'resnet18', 'squeezenet1_1', 'mobilenet_v3_small', 'vit_b_16', 'swin_t', 'efficientnet_b0',
        'convnext_tiny', 'densenet121', 'googlenet', 'regnet_y_400mf', 'shufflenet_v2_x0_5', 'mnasnet0_5',
        'swin_v2_t', 'vgg11_bn', 'resnet50']


def check(arch):
    rec = H.graph_records()[arch]
    g = Graph(H.build_model(arch), ve_cutoff=50, verbose=False)
    assert g.n_nodes == rec['n'], (arch, g.n_nodes, rec['n'])
    assert g.node_feat[:, 0].tolist() == rec['ops'], arch
    got = sorted(map(tuple, g.edges1.tolist()))
    assert got == sorted(map(tuple, rec['edges'])), arch
    info = [[[r[0], r[1], r[2], None if r[3] is None else list(r[3]), bool(r[4]), bool(r[5])] for r in cell]
            for cell in g.node_info]
    assert info == rec['node_info'], arch


@pytest.mark.parametrize('arch', FAST)
def test_tracer_matches_reference_graph(arch):
    check(arch)


@pytest.mark.slow
def test_tracer_matches_reference_graph_all():
    bad = []
    for arch in H.graph_records():
        if arch in FAST:
            continue
        try:
            check(arch)
        except AssertionError as e:
            bad.append((arch, str(e)[:80]))
    assert not bad, bad


def test_deepnets_style_generator_matches_reference_tracer():
    """ghn3_b200.deepnets.NetGenerator: deterministic per seed, every parameter mapped to a graph node, and our tracer
    reproduces the reference tracer's graphs of these cell networks (per-cell node_info) bit-exactly."""
    from ghn3_b200.deepnets import NetGenerator
    from ghn3_b200.plan import ModelPlan
    from ghn3_b200.weights import CONFIGS
    fx = H.cellnet_records()
    gen = NetGenerator(seed=fx['seed'])
    for name, rec in fx['graphs'].items():
        net = gen.sample_net()
        net.expected_input_sz = 64
        assert sum(p.numel() for p in net.parameters()) == rec['n_params'], name
        g = Graph(net, ve_cutoff=50, verbose=False)
        assert g.n_nodes == rec['n'] and g.node_feat[:, 0].tolist() == rec['ops'], name
        assert sorted(map(tuple, g.edges1.tolist())) == sorted(map(tuple, rec['edges'])), name
        info = [[[r[0], r[1], r[2], None if r[3] is None else list(r[3]), bool(r[4]), bool(r[5])] for r in cell]
                for cell in g.node_info]
        assert info == rec['node_info'], name
        assert len(g.node_info) == net.net_args['n_cells']
        plan = ModelPlan(g, net, CONFIGS['ghn3tm8'], True)
        assert plan.n_params == rec['n_params'], name
        assert net(torch.randn(2, 3, 64, 64)).shape == (2, 1000)


def test_msa_primitive_and_light_networks():
    """'msa' cell op (reference ghn3/ops.py:302): our tracer reproduces the reference tracer's graphs of msa-containing
    networks; the parameter-free twin (CellNetLight, the role of the reference's NetworkLight) has the same module
    names, so the same graph maps every node to a shape placeholder and plans the same prediction."""
    import gzip
    import json
    import os
    from ghn3_b200.deepnets import CellNetLight, NetGenerator, n_params_of
    from ghn3_b200.plan import BatchPlan, ModelPlan
    from ghn3_b200.weights import CONFIGS
    with gzip.open(os.path.join(H.GOLDEN, 'graphs_cellnets_msa.json.gz'), 'rt') as f:
        fx = json.load(f)
    gen = NetGenerator(seed=fx['seed'], with_msa=True, max_params=8e6)
    idx = -1
    for (name, rec), want in zip(fx['graphs'].items(), fx['stream_index']):
        while idx < want:
            net = gen.sample_net()
            idx += 1
        g_ = net.net_args['genotype']
        assert any(e[0] == 'msa' for e in g_['normal'] + g_['reduce'])
        net.expected_input_sz = 64
        g = Graph(net, ve_cutoff=50, verbose=False)
        assert g.n_nodes == rec['n'] and g.node_feat[:, 0].tolist() == rec['ops'], name
        assert sorted(map(tuple, g.edges1.tolist())) == sorted(map(tuple, rec['edges'])), name
        info = [[[r[0], r[1], r[2], None if r[3] is None else list(r[3]), bool(r[4]), bool(r[5])] for r in cell]
                for cell in g.node_info]
        assert info == rec['node_info'], name
        cfg = CONFIGS['ghn3tm8']
        plan = ModelPlan(g, net, cfg, True)
        assert plan.n_params == rec['n_params'] == n_params_of(net), name
        light = CellNetLight(**net.net_args)
        assert [n for n, _ in light.named_modules()] == [n for n, _ in net.named_modules()]
        assert list(light.parameters()) == [] and n_params_of(light) == rec['n_params']
        lplan = ModelPlan(g, light, cfg, True)
        assert (lplan.n_params, lplan.n_tensors) == (plan.n_params, plan.n_tensors)
        assert list(lplan.groups.items()) == list(plan.groups.items())
        a, b = BatchPlan([plan], cfg), BatchPlan([lplan], cfg)
        assert np.array_equal(a.desc_static, b.desc_static) and np.array_equal(a.dst_row, b.dst_row)
