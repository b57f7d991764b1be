"""cProfile of the host side of ghn(models, batch) and of the e2e step (run on the GPU box)."""
import cProfile, pstats, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from ghn3_b200 import GHN3, Graph, GraphBatch, param_norm
from ghn3_b200.weights import CONFIGS, procedural_state_dict
dev = torch.device('cuda:0')
cfg = CONFIGS['ghn3xlm16']
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
ghn.load_state_dict(procedural_state_dict(cfg, 0)); ghn = ghn.to(dev).eval()
records = bench.load_records()
models = [bench.build_model(a).to(dev) for a in bench.WORKLOAD_ARCHS]
graphs = [Graph.from_record(records[a]) for a in bench.WORKLOAD_ARCHS]
batch = GraphBatch(graphs, dense=True).to_device(dev)
with torch.no_grad():
    for _ in range(5): ghn(models, batch)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): ghn(models, batch)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print('host enqueue per forward: %.3f ms; incl. drain %.3f ms' % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(100): ghn(models, batch)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
    def e2e():
        b = GraphBatch(graphs, dense=True).to_device(dev)
        ghn(models, b)
        return [param_norm(m) for m in models]
    for _ in range(5): e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100): e2e()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print('e2e host enqueue per step: %.3f ms' % ((t1 - t0) / 100 * 1e3))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): GraphBatch(graphs, dense=True).to_device(dev)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
