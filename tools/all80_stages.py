"""Stage times (CUDA events between stages, serial) of ONE varlen batch with all 80 torchvision architectures."""
import gzip, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import GHN3, Graph, GraphBatch
from ghn3_b200.weights import CONFIGS, procedural_state_dict

name = sys.argv[1] if len(sys.argv) > 1 else 'ghn3xlm16'
dev = torch.device('cuda:0')
cfg = CONFIGS[name]
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
ghn.load_state_dict(procedural_state_dict(cfg, 0))
ghn = ghn.to(dev).eval()
records = bench.load_records()
names = sorted(records)
models = [bench.build_on_device(n, dev) for n in names]
graphs = [Graph.from_record(records[n]) for n in names]
batch = GraphBatch(graphs, dense=True).to_device(dev)
with torch.no_grad():
    for _ in range(3):
        ghn(models, batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ghn(models, batch)
    e1.record()
    torch.cuda.synchronize()
    print('%s all-80 batch: %.2f ms per call, %d nodes' % (name, e0.elapsed_time(e1) / 3, batch.pack.total_nodes))
    stage = {}
    for _ in range(3):
        ghn._profile = bench.PerOp()
        ghn(models, batch)
        torch.cuda.synchronize()
        ev = ghn._profile
        for (n0, a), (n1, b) in zip(ev[:-1], ev[1:]):
            k = n1.split('#')[0]
            stage[k] = stage.get(k, 0.0) + a.elapsed_time(b) / 3
    ghn._profile = None
    print({k: round(v, 3) for k, v in stage.items()})
