"""One Graphormer QKV GEMM launch (for `ncu --set full`): python tools/gemm_one.py [M] [mode: qkv|ffn1|ffn2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ghn3_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 18666
mode = sys.argv[2] if len(sys.argv) > 2 else 'qkv'
C = 384
dev = 'cuda'
h = torch.randn(M, C, device=dev).bfloat16()
if mode == 'qkv':
    w = (torch.randn(3 * C, C, device=dev) / 20).bfloat16()
    out = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(h, w, in_dtype=ops.BF16, out=out, out_dtype=ops.BF16, b_dynamic=False)
elif mode == 'ffn1':
    w = (torch.randn(4 * C, C, device=dev) / 20).bfloat16()
    b = torch.randn(4 * C, device=dev)
    out = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(h, w, bias=b, act=ops.ACT_GELU, in_dtype=ops.BF16, out=out, out_dtype=ops.BF16, b_dynamic=False)
else:
    ff = torch.randn(M, 4 * C, device=dev).bfloat16()
    w = (torch.randn(C, 4 * C, device=dev) / 40).bfloat16()
    b = torch.randn(C, device=dev)
    out = torch.randn(M, C, device=dev)
    fn = lambda: ops.gemm(ff, w, bias=b, in_dtype=ops.BF16, out=out, out_dtype=ops.F32, accumulate=True, b_dynamic=False)
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
