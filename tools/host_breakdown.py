import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch, bench
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
def T(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e3
with torch.no_grad():
    ghn(models, batch)
    w = ghn._device_weights()
    bp = ghn._batch_plan(batch, models, True, False)
    prog = bp.program
    print('weights_signature   %.3f ms' % T(lambda: ghn._weights_signature()))
    print('batch_plan lookup   %.3f ms' % T(lambda: ghn._batch_plan(batch, models, True, False)))
    print('refresh_targets     %.3f ms' % T(lambda: prog.refresh_targets(True)))
    print('tok normal_         %.3f ms' % T(lambda: prog.tok.normal_(mean=0.0, std=0.02)))
    print('run (C sequence)    %.3f ms' % T(lambda: prog.run(None), 100))
    print('full forward        %.3f ms' % T(lambda: ghn(models, batch), 100))
    print('GraphBatch+to_dev   %.3f ms' % T(lambda: GraphBatch(graphs, dense=True).to_device(dev)))
    print('param_norm x2       %.3f ms' % T(lambda: [param_norm(m) for m in models]))
    # true host cost of the C sequence: GPU idle before every call, do not wait for completion
    import time as _t
    tt = 0.0
    for _ in range(50):
        torch.cuda.synchronize(); t0 = _t.perf_counter(); prog.run(None); tt += _t.perf_counter() - t0
    torch.cuda.synchronize()
    print('run() host-only     %.3f ms' % (tt / 50 * 1e3))
    x = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    for name, fn in (('memset 1GiB', lambda: x.zero_()), ('copy 0.5GiB', lambda: x[:1 << 29].copy_(x[1 << 29:]))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print('%s: %.3f ms -> %.0f GB/s' % (name, ms, (1 << 30) / ms / 1e6))
