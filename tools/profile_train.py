"""Where a training step spends its time: host wall clock per phase + device time per backward op."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import GHN3, Graph, GraphBatch, Trainer
from ghn3_b200.weights import CONFIGS, procedural_state_dict

ap = argparse.ArgumentParser()
ap.add_argument('--cfg', default='ghn3xlm16')
ap.add_argument('--dtype', default='bf16')
ap.add_argument('--archs', default='netgen', help="'netgen' = NetGenerator(0).sample(8), else comma-separated torchvision names")
a = ap.parse_args()
dev = torch.device('cuda')
cfg = CONFIGS[a.cfg]
records = bench.load_records()
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=a.dtype)
ghn.load_state_dict(procedural_state_dict(cfg, 0))
ghn = ghn.to(dev).train()
if a.archs == 'netgen':
    from ghn3_b200.deepnets import NetGenerator
    pairs = NetGenerator(seed=0).sample(8)
    graphs = GraphBatch([g for _, g in pairs], dense=True).to_device(dev)
    nets = [n.to(dev) for n, _ in pairs]
else:
    archs = a.archs.split(',')
    graphs = GraphBatch([Graph.from_record(records[x]) for x in archs], dense=True).to_device(dev)
    nets = [bench.build_model(x).to(dev) for x in archs]
opt = torch.optim.AdamW(ghn.parameters(), lr=4e-4, weight_decay=1e-2)


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def step(profile=False):
    t = {}
    t0 = sync()
    opt.zero_grad(set_to_none=True)
    out = ghn(nets, graphs, keep_grads=True, reduce_graph=True)
    t1 = sync(); t['forward (incl. weight conversion)'] = t1 - t0
    loss = ghn.last_program.pred_flat.sum() * 1e-3
    t2 = sync(); t['stub loss'] = t2 - t1
    if profile:
        ghn._profile_bwd = []
    loss.backward()
    t3 = sync(); t['backward'] = t3 - t2
    torch.nn.utils.clip_grad_norm_(ghn.parameters(), 5)
    t4 = sync(); t['clip'] = t4 - t3
    opt.step()
    t5 = sync(); t['adamw'] = t5 - t4
    return t


for _ in range(2):
    step()
t = step()
print('--- host wall clock per phase (ms), synchronised ---')
for k, v in t.items():
    print('%-40s %9.2f' % (k, v * 1e3))
t = step(profile=True)
ev = ghn._profile_bwd
ghn._profile_bwd = None
agg = {}
for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
    agg.setdefault(n1, [0.0, 0])
    agg[n1][0] += e0.elapsed_time(e1)
    agg[n1][1] += 1
print('--- backward ops (device ms, count) ---')
for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-24s %9.3f  x%d' % (k, ms, n))
bp = ghn.last_program.bp
print('segments', len(bp.segments), 'rows', bp.conv_total_rows, 'nodes', bp.total_nodes, 'fc problems', len(bp.fc_problems),
      'descs', len(bp.desc_static))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
