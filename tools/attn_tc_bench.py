"""tcgen05 attention vs the mma.sync kernel: ms per layer at ghn3xlm16 (16 heads x 24) on synthetic DAGs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ghn3_b200 import _lib as L, ops
from ghn3_b200.synthetic import synthetic_dag

C, H = 384, 16
lib = L.load()
for n in (152, 305, 512, 816, 1024, 2048, 4096):
    e, op = synthetic_dag(n, 0)
    pack = ops.GraphPack([n], edges=[e], cutoff=50, device='cuda', op=op).build()
    qkv = torch.randn(n, 3 * C, device='cuda').bfloat16()
    lut = torch.randn(H, 51 * 51, device='cuda')
    res = {}
    for name, thr in (('mma.sync', 1 << 30), ('tcgen05', 0)):
        lib.ghn3_set_attention_tc_min(thr)
        for _ in range(3):
            ops.attention(qkv, pack, lut, C, H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention(qkv, pack, lut, C, H)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20
    fl = 4.0 * n * n * C
    print('N=%5d  mma.sync %.4f ms (%.1f TF/s)   tcgen05 %.4f ms (%.1f TF/s)   x%.2f' % (
        n, res['mma.sync'], fl / res['mma.sync'] / 1e9, res['tcgen05'], fl / res['tcgen05'] / 1e9,
        res['mma.sync'] / res['tcgen05']))
lib.ghn3_set_attention_tc_min(2048)
