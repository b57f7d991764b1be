"""Times ghn3_graphormer_fused against ghn3_graphormer_stack on the bench workload (ViT-B/16 + ConvNeXt-Base)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_fused_gpu import _setup

C, Hh, layers = 384, 16, 24
archs = sys.argv[1].split(',') if len(sys.argv) > 1 else ['vit_b_16', 'convnext_base']
recs, pack, per, stack, x0, lut, fg = _setup(C, Hh, layers, archs, seed=3)
for stop in (0, 1, 2, 3, 4, 5, 10, 60):
    for _ in range(3):
        fg.run(stop_after=stop)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fg.run(stop_after=stop)
    e1.record()
    torch.cuda.synchronize()
    print('stop_after=%3d  nodes=%d  %.1f us per call' % (stop, pack.total_nodes, e0.elapsed_time(e1) / 20 * 1e3))
