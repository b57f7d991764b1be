"""cProfile of COLD predictions: one ghn(model, graph) call per fresh architecture (plans, program build, uploads)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import GHN3, Graph
from ghn3_b200.weights import CONFIGS, procedural_state_dict

dev = torch.device('cuda:0')
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'ghn3xlm16']
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
ghn.load_state_dict(procedural_state_dict(cfg, 0))
ghn = ghn.to(dev).eval()
records = bench.load_records()
names = sorted(records)[:40]
models = [bench.build_on_device(n, dev) for n in names]
graphs = [Graph.from_record(records[n]) for n in names]
with torch.no_grad():
    ghn(bench.build_on_device('resnet18', dev), Graph.from_record(records['resnet18']))      # CUDA / LUT / kernels warm
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for m, g in zip(models, graphs):
        ghn(m, g)
    pr.disable()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print('cold: %.2f ms per model on the host (%.2f incl. device drain)' % ((t1 - t0) / len(names) * 1e3, (t2 - t0) / len(names) * 1e3))
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
