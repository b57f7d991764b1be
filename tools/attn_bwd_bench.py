"""Times ghn3_attention_bwd (XL head geometry) on the training meta-batch graphs, with and without the LUT gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ghn3_b200 import ops
from ghn3_b200 import GraphBatch
from ghn3_b200.deepnets import NetGenerator

dev = torch.device('cuda')
pairs = NetGenerator(seed=0).sample(8)
batch = GraphBatch([g for _, g in pairs], dense=True).to_device(dev)
pack = batch.pack
C, H = 384, 16
N = pack.total_nodes
torch.manual_seed(0)
for dtype in (torch.bfloat16, torch.float32):
    qkv = torch.randn(N, 3 * C, device=dev).to(dtype)
    d_out = torch.randn(N, C, device=dev).to(dtype)
    lut = torch.randn(H, 51 * 51, device=dev) * 0.5
    out = ops.attention(qkv, pack, lut, C, H, dtype=ops.BF16 if dtype == torch.bfloat16 else ops.F32)
    lse2 = torch.zeros(H, N, device=dev)
    if dtype == torch.bfloat16:
        out = ops.attention(qkv, pack, lut, C, H, dtype=ops.BF16, lse2=lse2)
    for mma in ((False, True) if dtype == torch.bfloat16 else (False,)):
        for with_lut in (True, False):
            d_lut = torch.zeros_like(lut) if with_lut else None
            kw = dict(d_lut=d_lut, fwd_lse2=lse2 if mma else None)
            for _ in range(3):
                ops.attention_bwd(qkv, out, d_out, pack, lut, C, H, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.attention_bwd(qkv, out, d_out, pack, lut, C, H, **kw)
            e1.record()
            torch.cuda.synchronize()
            print('%s mma=%-5s d_lut=%-5s  %.1f us per call (incl. scratch alloc + binning when d_lut) nodes %s' % (
                dtype, mma, with_lut, e0.elapsed_time(e1) / 20 * 1e3, [g.n_nodes for _, g in pairs]))
