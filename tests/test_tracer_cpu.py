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
