"""Times ghn3_attention (XL head geometry) on the bench workload graphs: bf16 MMA kernel vs the fp32-storage kernels
(split-bf16 tensor-core kernel; CUDA-core kernel with GHN3_NO_SPLIT_ATTN=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import ops, Graph, GraphBatch

dev = torch.device('cuda')
records = bench.load_records()
if len(sys.argv) > 1 and sys.argv[1] == 'netgen':          # the training meta-batch: 8 graphs, 100-507 nodes
    from ghn3_b200.deepnets import NetGenerator
    batch = GraphBatch([g for _, g in NetGenerator(seed=0).sample(8)], dense=True).to_device(dev)
else:
    batch = GraphBatch([Graph.from_record(records[a]) for a in bench.WORKLOAD_ARCHS], dense=True).to_device(dev)
pack = batch.pack
C, H = 384, 16
N = pack.total_nodes
torch.manual_seed(0)
qkv = torch.randn(N, 3 * C, device=dev)
lut = torch.randn(H, 51 * 51, device=dev) * 0.5
ref = None
for name, x, dt in (('bf16', qkv.bfloat16(), ops.BF16), ('fp32 storage', qkv, ops.F32)):
    for _ in range(5):
        out = ops.attention(x, pack, lut, C, H, dtype=dt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        out = ops.attention(x, pack, lut, C, H, dtype=dt)
    e1.record()
    torch.cuda.synchronize()
    print('%-14s %.1f us per call (N=%d, split kernel %s)' % (name, e0.elapsed_time(e1) / 50 * 1e3, N,
                                                            'off' if os.environ.get('GHN3_NO_SPLIT_ATTN') else 'on'))
