"""Runs a few steps of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import argparse, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from ghn3_b200 import GHN3, Graph, GraphBatch
from ghn3_b200.weights import CONFIGS, procedural_state_dict

ap = argparse.ArgumentParser()
ap.add_argument('--dtype', default='bf16')
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--cfg', default='ghn3xlm16')
ap.add_argument('--archs', default=','.join(bench.WORKLOAD_ARCHS))
ap.add_argument('--stage', default='', help='profile only this stage of the step (e.g. dec_conv2, scatter, graphormer)')
args = ap.parse_args()
dev = torch.device('cuda:0')
cfg = CONFIGS[args.cfg]
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=args.dtype)
ghn.load_state_dict(procedural_state_dict(cfg, 0))
ghn = ghn.to(dev).eval()
records = bench.load_records()
archs = args.archs.split(',')
models = [bench.build_model(a).to(dev) for a in archs]
graphs = [Graph.from_record(records[a]) for a in archs]
batch = GraphBatch(graphs, dense=True).to_device(dev)
with torch.no_grad():
    for _ in range(args.warmup):
        ghn(models, batch)
    torch.cuda.synchronize()
    if args.stage:
        order = ['start', 'node_features', 'graphormer', 'dec_fc', 'dec_conv0', 'dec_conv2', 'heads_1d', 'scatter']
        prev = order[order.index(args.stage) - 1]

        class Hook(list):
            def append(self, item):
                if item[0] == prev:
                    torch.cuda.synchronize(); torch.cuda.profiler.start()
                if item[0] == args.stage:
                    torch.cuda.synchronize(); torch.cuda.profiler.stop()
        ghn._profile = Hook()
        for _ in range(args.steps):
            ghn(models, batch)
        torch.cuda.synchronize()
    else:
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            ghn(models, batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
print('profiled', args.steps, 'steps')
