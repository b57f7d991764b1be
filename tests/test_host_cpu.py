"""CPU-only tests: host planning logic against the oracle, state_dict contract, C-ABI exports, sharding (gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from ghn3_b200.graph import Graph, GraphBatch
from ghn3_b200.plan import BatchPlan, ModelPlan, ShapeIndexer, layered_modules, scale_for
from ghn3_b200.weights import CONFIGS, procedural_state_dict, state_dict_spec
from oracle import ghn3_oracle as O
from tests import helpers as H

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shape_indexer_matches_oracle_tables():
    for ncls, ms in [(1000, (384, 384, 16, 16)), (10, (64, 64, 11, 11)), (1000, (64, 64, 16, 16))]:
        idx, tabs = ShapeIndexer(ncls, ms), O.ShapeTables(ncls, ms)
        assert tuple(idx.dummy) == tabs.dummy
        for v in list(range(1, 700)) + [1000, 1024, 4095, 4096, 7392, 8192, 8193, 10000, 25088]:
            for sz in [(v,), (v, 3), (3, v), (64, v, 7, 7), (v, 1, 3, 3), (1, 197, v)]:
                assert tuple(idx.lookup(sz)) == tabs.indices(sz), (sz,)
        for k in list(range(1, 40)) + [64, 224]:
            assert tuple(idx.lookup((8, 8, k, k))) == tabs.indices((8, 8, k, k))


ARCHS = ['resnet50', 'vit_b_16', 'convnext_base', 'squeezenet1_1', 'mobilenet_v3_small', 'swin_v2_t', 'densenet121',
         'inception_v3', 'vit_b_32', 'efficientnet_b0', 'regnet_y_400mf', 'alexnet']


@pytest.mark.parametrize('cfg_name', ['ghn3tm8', 'ghn3xlm16'])
@pytest.mark.parametrize('arch', ARCHS)
def test_plan_matches_oracle_mapping(arch, cfg_name):
    cfg = CONFIGS[cfg_name]
    rec = H.graph_records()[arch]
    model = H.build_model(arch)
    g = Graph.from_record(rec)
    plan = ModelPlan(g, model, cfg)
    og = O.graph_from_record(rec, cutoff=1)
    mapping, params_map = O.map_net_params(og['node_info'], model, cfg['max_shape'])
    # group keys and member order: bit-exact (nn.py:677-680)
    assert list(plan.groups.keys()) == list(mapping.keys())
    for k in mapping:
        assert plan.groups[k] == mapping[k], k
    assert np.array_equal(plan.shape_idx, O.node_shape_indices(og['node_info'], model, cfg, rec['n']))
    # every predicted tensor is covered exactly once and the parameter count matches the reference's self-check
    bp = BatchPlan([plan], cfg)
    covered = {}
    for (module, attr, shape, view), d in zip(bp.desc_targets, bp.desc_static):
        key = (id(module), attr)
        covered[key] = covered.get(key, 0) + int(d['numel'])
    total = 0
    for key, n in covered.items():
        total += n
    assert total == plan.n_params
    expected = 0
    for node, (matched, key, pos) in params_map.items():
        if pos is None:
            continue
        sz = matched['sz']
        expected += int(np.prod(sz)) * (2 if (len(sz) == 1 and matched['is_w']) else 1)
    assert plan.n_params == expected


def test_resnet50_counts_and_groups():
    cfg = CONFIGS['ghn3xlm16']
    plan = ModelPlan(Graph.from_record(H.graph_records()['resnet50']), H.build_model('resnet50'), cfg)
    assert (plan.n_tensors, plan.n_params) == (161, 25557032)          # SURVEY.md §4 / nn.py:384-392
    # the 15 shape groups of ResNet-50 at XL
    expect = [((64, 4, 7, 7), 1), ((64, 0), 7), ((64, 64, 1, 1), 1), ((384, 64, 1, 1), 4), ((384, 0), 38),
              ((64, 64, 3, 3), 3), ((64, 384, 1, 1), 2), ((128, 384, 1, 1), 4), ((384, 384, 1, 1), 21),
              ((128, 0), 8), ((128, 128, 3, 3), 4), ((384, 128, 1, 1), 4), ((384, 384, 3, 3), 9), ((384, 384), 1),
              ((384, -1), 1)]
    assert [(k, len(v)) for k, v in plan.groups.items()] == expect


def test_batch_plan_conv2_problems_cover_needed_columns_only():
    cfg = CONFIGS['ghn3xlm16']
    plan = ModelPlan(Graph.from_record(H.graph_records()['resnet50']), H.build_model('resnet50'), cfg)
    bp = BatchPlan([plan], cfg)
    ms1 = cfg['max_shape'][1]
    rows = sum(r[1] for r in bp.conv_rows)
    assert rows == bp.conv_total_rows
    # every conv node's (o', i') block is produced by exactly the problems of its class
    for q, (b, t) in enumerate(bp.conv):
        r0, P, _, _ = bp.conv_rows[q]
        hits = [p for _, pr, _ in bp.c2_launches for p in pr if p['a_row0'] <= r0 < p['a_row0'] + p['m']]
        cols = sum(int(p['n']) for p in hits)
        assert cols == t.o_need * t.i_need, (t.key, cols)
    assert bp.dst_row.max() == bp.n_conv + bp.n_1d - 1
    assert len(set(bp.dst_row[bp.dst_row >= 0])) == bp.n_conv + bp.n_1d


def test_scale_for_matches_oracle_normalize():
    for sz in [(64, 3, 7, 7), (96, 1, 5, 5), (32, 16, 1, 3), (1000, 2048), (1, 197, 768), (1, 8, 11, 11), (8, 8, 3, 3)]:
        p = torch.ones(sz)
        ref = O.normalize(p, True)
        assert abs(ref.flatten()[0].item() - scale_for(sz)) < 1e-9, sz


def test_state_dict_contract_and_from_pretrained(tmp_path):
    from ghn3_b200.nn import GHN3, from_pretrained
    cfg = CONFIGS['ghn3tiny']
    ghn = GHN3(**cfg, weight_norm=True, ve=True)
    spec = state_dict_spec(cfg)
    got = {k: tuple(v.shape) for k, v in ghn.state_dict().items()}
    assert got == {k: tuple(v) for k, v in spec.items()}
    sd = procedural_state_dict(cfg, 0)
    ghn.load_state_dict(sd)
    # checkpoint layouts of the reference: {'state_dict', 'config', ...} (trainer.py:419-426) and a bare state_dict
    path = tmp_path / 'ckpt.pt'
    torch.save({'state_dict': sd, 'config': dict(cfg, weight_norm=True, ve=True), 'epoch': 0, 'step': 0}, path)
    g2 = from_pretrained(str(path))
    assert all(torch.equal(a, b) for a, b in zip(g2.state_dict().values(), ghn.state_dict().values()))
    path2 = tmp_path / 'bare.pt'
    torch.save({'state_dict': sd}, path2)
    g3 = from_pretrained(str(path2))                 # config inferred from names/shapes (nn.py:59-100)
    assert (g3.hid, g3.layers, g3.heads, g3.max_shape, g3.num_classes) == (32, 2, 8, (32, 32, 16, 16), 1000)
    # pretrained=True layout: structural embeddings at top level until fix_embed_layers (nn.py:156-158,174-184)
    g4 = GHN3(**cfg, weight_norm=True, ve=True, pretrained=True)
    assert 'centrality_embed_in.weight' in g4.state_dict()
    g4.fix_embed_layers()
    assert 'gnn.0.centrality_embed_in.weight' in g4.state_dict()
    with pytest.raises(NotImplementedError):
        GHN3(**cfg, is_ghn2=True)
    # the released checkpoints are bare state_dicts pickled with joblib (reference nn.py:49)
    import joblib
    path3 = tmp_path / 'ghn3_joblib.pt'
    joblib.dump(sd, path3)
    g5 = from_pretrained(str(path3))
    assert (g5.hid, g5.layers, g5.layernorm) == (32, 2, True)
    assert all(torch.equal(a, b) for a, b in zip(g5.state_dict().values(), ghn.state_dict().values()))
    # config inference over every shipped configuration, including layernorm=False / 11x11 decoder grids
    from ghn3_b200.nn import infer_config
    for name, c in CONFIGS.items():
        got = infer_config(procedural_state_dict(c, 0))
        assert (got['hid'], got['layers'], got['max_shape'], got['num_classes'], got['layernorm']) == \
            (c['hid'], c['layers'], tuple(c['max_shape']), c['num_classes'], True), name
    c10 = dict(hid=32, layers=1, heads=8, max_shape=(32, 32, 11, 11), num_classes=10, layernorm=False)
    got = infer_config(procedural_state_dict(c10, 0))
    assert (got['max_shape'], got['num_classes'], got['layernorm'], got['layers']) == ((32, 32, 11, 11), 10, False, 1)


def test_no_cpu_fallback():
    from ghn3_b200.nn import GHN3
    cfg = CONFIGS['ghn3tiny']
    ghn = GHN3(**cfg, weight_norm=True, ve=True)
    with pytest.raises(RuntimeError, match='CUDA'):
        ghn(H.build_model('resnet18'), Graph.from_record(H.graph_records()['resnet18']))
    with pytest.raises(RuntimeError, match='CUDA'):
        GraphBatch([Graph.from_record(H.graph_records()['resnet18'])]).to_device('cpu')


def test_c_abi_exports_every_declared_symbol():
    from ghn3_b200 import _lib
    header = open(os.path.join(REPO, 'include', 'ghn3_b200.h')).read()
    declared = set(re.findall(r'\b(ghn3_[a-z0-9_]+)\s*\(', header))
    declared -= {'ghn3_stream_t'}
    lib = _lib.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert set(_lib.SYMBOLS_ALL) == declared
    assert lib.ghn3_abi_version() == int(re.search(r'#define GHN3_ABI_VERSION (\d+)', header).group(1))
    assert lib.ghn3_launch_count() == 0


def test_graph_containers_host_logic():
    recs = H.graph_records()
    g1, g2 = Graph.from_record(recs['resnet18']), Graph.from_record(recs['alexnet'])
    b = GraphBatch([g1, g2], dense=True)
    assert len(b) == 2 and b.n_nodes == [53, 21] and b[1] is g2 and list(b) == [g1, g2]
    x = torch.arange(74.).view(74, 1)
    dense, off = b.to_dense(x)
    assert dense.shape == (2, 53, 1) and off == [0, 53, 74]
    assert torch.equal(b.to_sparse(dense), x)
    # reference-style construction from a dense adjacency with virtual edges
    og = O.graph_from_record(recs['alexnet'])
    gd = Graph(node_feat=torch.as_tensor(og['ops']).view(-1, 1), node_info=og['node_info'], A=og['A'], dense=True)
    assert np.array_equal(gd.edges1, g2.edges1[np.lexsort((g2.edges1[:, 1], g2.edges1[:, 0]))])
    assert torch.equal(gd._Adj, torch.as_tensor(og['A']))


def test_lpt_sharding():
    from ghn3_b200.shard import shard_lpt
    costs = [r['n_params'] for r in H.graph_records().values()]
    for world in (1, 2, 4, 8):
        shards = shard_lpt(costs, world)
        assert sorted(sum(shards, [])) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= sum(costs) / world + max(costs)


def test_sharding_two_ranks_gloo(tmp_path):
    """world_size-2 run of the N>1 inference path's host logic: disjoint cover, no data-path collective."""
    script = tmp_path / 'w.py'
    script.write_text('''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from ghn3_b200.shard import my_shard
from tests import helpers as H
dist.init_process_group('gloo')
costs = [r['n_params'] for r in H.graph_records().values()]
mine = my_shard(costs)
out = [None, None]
dist.all_gather_object(out, mine)
if dist.get_rank() == 0:
    assert sorted(out[0] + out[1]) == list(range(len(costs))) and not set(out[0]) & set(out[1])
    print('OK', len(out[0]), len(out[1]))
dist.destroy_process_group()
''' % REPO)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29531', str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout + r.stderr


def test_training_exchange_two_ranks_gloo(tmp_path):
    """world_size-2 run of the training path's host logic: contiguous meta-batch shares (train_ghn_ddp.py:92) and the
    two-region flat-gradient average of ghn3_b200.train.GradSync (sum + divide on gloo, AVG on NCCL)."""
    script = tmp_path / 'w.py'
    script.write_text('''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from ghn3_b200.train import GradSync
from ghn3_b200.trainer import shard_meta_batch
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard_meta_batch(8, rank, world)
assert mine == list(range(rank * 4, rank * 4 + 4))
try:
    shard_meta_batch(7, rank, world)
    raise SystemExit('expected ValueError')
except ValueError:
    pass
sync = GradSync()
assert not sync.avg and sync.world == 2
flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
works = [sync.start(flat[:6]), sync.start(flat[6:])]
sync.finish(works, flat)
assert torch.allclose(flat, torch.arange(10, dtype=torch.float32) * 1.5), flat
if rank == 0:
    print('OK')
dist.destroy_process_group()
''' % REPO)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29532', str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout + r.stderr


def test_flat_gradient_layout_and_conv2_block_lists():
    """Host logic of the training path: [decoder | rest] gradient layout with 16-byte aligned slices, and the sparse-K
    block lists of the conv.2 backward GEMMs cover every non-zero block of the column-expanded operands."""
    from ghn3_b200.nn import GHN3
    from ghn3_b200.train import flat_layout, _conv2_k_blocks
    cfg = CONFIGS['ghn3tiny']
    ghn = GHN3(**cfg, weight_norm=True, ve=True)
    params, order, offs, early = flat_layout(ghn)
    assert len(order) == len(params) == len(list(ghn.state_dict())) and {id(p) for p in order} == {id(p) for p in params}
    assert all(int(o) % 4 == 0 for o in offs) and int(offs[-1]) >= sum(p.numel() for p in params)
    n_early = sum(p.numel() for m in (ghn.decoder, ghn.decoder_1d, ghn.bias_class) for p in m.parameters())
    assert early >= n_early and early < int(offs[-1])
    dec_ids = {id(p) for m in (ghn.decoder, ghn.decoder_1d, ghn.bias_class) for p in m.parameters()}
    k = len(dec_ids)
    assert {id(p) for p in order[:k]} == dec_ids                      # decoder-side tensors first

    ms1, KF = 24, 24 * 24
    segs = [(20, 24, 0, 130, 0), (7, 5, 130, 9, 0), (3, 1, 139, 200, 0)]
    R = 339
    X = np.zeros((R, KF), bool)
    for (o, ii, r0, rows, _) in segs:
        for a in range(o):
            X[r0:r0 + rows, a * ms1:a * ms1 + ii] = True
    for bk in (64, 32):
        (dl, do), (wl, wo) = _conv2_k_blocks(segs, R, KF, ms1, bk, 'cpu')
        for mt in range(-(-R // 128)):
            need = set((np.nonzero(X[mt * 128:(mt + 1) * 128].any(0))[0] // bk).tolist())
            assert need <= set(dl[do[mt]:do[mt + 1]].tolist())
        for ct_ in range(-(-KF // 128)):
            need = set((np.nonzero(X[:, ct_ * 128:(ct_ + 1) * 128].any(1))[0] // bk).tolist())
            assert need <= set(wl[wo[ct_]:wo[ct_ + 1]].tolist())
        assert len(dl) < (-(-R // 128)) * (KF // bk)                  # and they do skip something


def test_flat_layout_regions_split_into_equal_aligned_shards():
    """Sharded optimizer step: both regions of the flat gradient / parameter layout are multiples of FLAT_PAD elements,
    so every world size dividing FLAT_PAD / 4 gets equal 16-byte aligned slices; the chunk range of a slice found by
    bisection covers exactly the chunks that intersect it."""
    from ghn3_b200.nn import GHN3
    from ghn3_b200.train import FLAT_PAD, flat_layout
    ghn = GHN3(**CONFIGS['ghn3tiny'], weight_norm=True, ve=True)
    params, order, offs, early = flat_layout(ghn)
    total = int(offs[-1])
    assert early % FLAT_PAD == 0 and total % FLAT_PAD == 0 and 0 < early < total
    assert all(int(offs[i]) + order[i].numel() <= int(offs[i + 1]) for i in range(len(order)))
    CH = 8192
    numels = np.array([p.numel() for p in order], dtype=np.int64)
    chunks = (numels + CH - 1) // CH
    chunk0 = np.concatenate([[0], np.cumsum(chunks)[:-1]])
    start = np.repeat(offs[:-1], chunks) + (np.arange(int(chunks.sum())) - np.repeat(chunk0, chunks)) * CH
    end = np.minimum(start + CH, np.repeat(offs[:-1] + numels, chunks))
    for world in (2, 4, 8):
        for lo_r, hi_r in ((0, early), (early, total)):
            n = (hi_r - lo_r) // world
            assert n % 4 == 0
            for r in range(world):
                lo, hi = lo_r + r * n, lo_r + (r + 1) * n
                c0 = max(int(np.searchsorted(start, lo, side='right')) - 1, 0)
                c1 = int(np.searchsorted(start, hi, side='left'))
                hit = np.nonzero((end > lo) & (start < hi))[0]
                assert len(hit) == 0 or (c0 <= hit[0] and hit[-1] < c1)


def test_vectorised_plan_helpers_match_their_scalar_definitions():
    """tiles_for (one vectorised pass over all problems) against the per-problem enumeration; fastdiv_array against the
    scalar magic-number routine the scatter descriptors were defined with."""
    from ghn3_b200._lib import fastdiv
    from ghn3_b200.plan import PROBLEM_DT, fastdiv_array, tiles_for
    rs = np.random.RandomState(0)
    probs = np.zeros(37, dtype=PROBLEM_DT)
    probs['m'] = rs.randint(0, 700, 37)
    probs['n'] = rs.randint(0, 1500, 37)
    for bm, bn in ((128, 128), (64, 128), (128, 120)):
        want = [(p, i, j, 0) for p in range(37) for i in range(-(-int(probs['m'][p]) // bm))
                for j in range(-(-int(probs['n'][p]) // bn))]
        got = tiles_for(probs, block_m=bm, block_n=bn)
        assert got.dtype == np.int32 and got.tolist() == [list(t) for t in want]
    assert tiles_for(probs[:0]).shape == (0, 4)
    ds = np.concatenate([np.arange(0, 3000), 2 ** np.arange(1, 31), 2 ** np.arange(1, 31) - 1, 2 ** np.arange(1, 31) + 1,
                         rs.randint(1, 2 ** 31 - 1, 5000)])
    mul, sh = fastdiv_array(ds)
    assert all((int(a), int(b)) == fastdiv(int(d)) for a, b, d in zip(mul, sh, ds))
    n = rs.randint(0, 2 ** 31 - 1, 2000).astype(np.uint64)
    for d in (3, 7, 24, 384, 1000, 65537):
        m_, s_ = fastdiv(d)
        assert np.array_equal((n * np.uint64(m_) >> np.uint64(32)) >> np.uint64(s_), n // np.uint64(d))


def test_net_generator_is_deterministic_and_within_the_design_space():
    from ghn3_b200.deepnets import NetGenerator, CHANNELS, OPS
    a, b = NetGenerator(seed=3), NetGenerator(seed=3)
    for _ in range(4):
        na, nb = a.sample_net(), b.sample_net()
        assert na.net_args == nb.net_args
        g = na.net_args
        assert 4 <= g['n_cells'] <= 18 and g['C'] in CHANNELS and g['stem_type'] in (0, 1) and g['fc_layers'] in (1, 2)
        assert all(op in OPS and op != 'none' for cell in ('normal', 'reduce') for (op, _, _) in g['genotype'][cell])
        assert sum(p.numel() for p in na.parameters()) <= a.max_params
    assert NetGenerator(seed=4).sample_net().net_args != NetGenerator(seed=3).sample_net().net_args


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on the host cores only and prints ONE JSON line with the contract's keys."""
    import json
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'config', 'e2e',
              'cpu_baseline', 'impl'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'models/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']
