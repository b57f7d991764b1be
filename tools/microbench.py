"""Steady-state per-launch time of the small Graphormer kernels (back-to-back launches on one stream)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ghn3_b200 import ops

def timeit(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

dev = 'cuda'
M, C = 457, 384
x = torch.randn(M, C, device=dev)
h = torch.randn(M, C, device=dev).bfloat16()
ff = torch.randn(M, 4 * C, device=dev).bfloat16()
# 24 distinct weight sets so that weights are cold like in the real stack (170 MB > L2? no: 24*3.5MB = 85 MB; use 48)
L = 48
wq = [(torch.randn(3 * C, C, device=dev) / 20).bfloat16() for _ in range(L)]
wo = [(torch.randn(C, C, device=dev) / 20).bfloat16() for _ in range(L)]
w1 = [(torch.randn(4 * C, C, device=dev) / 20).bfloat16() for _ in range(L)]
w2 = [(torch.randn(C, 4 * C, device=dev) / 40).bfloat16() for _ in range(L)]
b = torch.randn(4 * C, device=dev)
qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
f1 = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
g = torch.ones(C, device=dev); bb = torch.zeros(C, device=dev)
i = [0]
def nxt():
    i[0] = (i[0] + 1) % L
    return i[0]
print('layernorm            %.2f us' % timeit(lambda: ops.layernorm(x, g, bb, out_dtype=ops.BF16)))
for bn in (64, 128):
    print('QKV  bn=%3d          %.2f us' % (bn, timeit(lambda: ops.gemm(h, wq[nxt()], in_dtype=ops.BF16, out=qkv, out_dtype=ops.BF16, block_n=bn))))
    print('FF1  bn=%3d gelu     %.2f us' % (bn, timeit(lambda: ops.gemm(h, w1[nxt()], bias=b, act=ops.ACT_GELU, in_dtype=ops.BF16, out=f1, out_dtype=ops.BF16, block_n=bn))))
    for sp in (1, 3, 6):
        print('proj bn=%3d split=%d  %.2f us' % (bn, sp, timeit(lambda: ops.gemm(h, wo[nxt()], bias=b, in_dtype=ops.BF16, out=x, out_dtype=ops.F32, accumulate=True, block_n=bn, k_splits=sp))))
        print('FF2  bn=%3d split=%d  %.2f us' % (bn, sp, timeit(lambda: ops.gemm(ff, w2[nxt()], bias=b, in_dtype=ops.BF16, out=x, out_dtype=ops.F32, accumulate=True, block_n=bn, k_splits=sp))))
# same weights every time (L2-warm) for comparison
print('QKV warm weights     %.2f us' % timeit(lambda: ops.gemm(h, wq[0], in_dtype=ops.BF16, out=qkv, out_dtype=ops.BF16)))
print('FF2 warm split auto  %.2f us' % timeit(lambda: ops.gemm(ff, w2[0], bias=b, in_dtype=ops.BF16, out=x, out_dtype=ops.F32, accumulate=True)))
# torch reference points
hw = wq[0]
print('torch F.linear QKV   %.2f us' % timeit(lambda: torch.nn.functional.linear(h, wq[nxt()])))
print('torch F.linear FF2   %.2f us' % timeit(lambda: torch.nn.functional.linear(ff, w2[nxt()])))
print('torch empty kernel   %.2f us' % timeit(lambda: x.add_(0.0)))
