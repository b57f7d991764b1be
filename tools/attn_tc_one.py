import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ghn3_b200 import _lib as L, ops
from ghn3_b200.synthetic import synthetic_dag
C, H, n = 384, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
e, op = synthetic_dag(n, 0)
pack = ops.GraphPack([n], edges=[e], cutoff=50, device='cuda', op=op).build()
qkv = torch.randn(n, 3 * C, device='cuda').bfloat16()
lut = torch.randn(H, 51 * 51, device='cuda')
L.load().ghn3_set_attention_tc_min(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
for _ in range(3):
    ops.attention(qkv, pack, lut, C, H)
torch.cuda.synchronize()
