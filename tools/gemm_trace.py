"""Per-CTA phase timeline of one GEMM launch (globaltimer stamps written by the kernel)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from ghn3_b200 import ops, _lib as L
dev='cuda'
lib = L.load()
buf = torch.zeros(4096, 8, dtype=torch.int64, device=dev)
def trace(name, fn, nctas):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    # flush L2 so weights are cold like in the real step
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev); junk.fill_(1); torch.cuda.synchronize()
    buf.zero_(); lib.ghn3_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
    fn(); torch.cuda.synchronize()
    lib.ghn3_debug_gemm_trace(ctypes.c_void_p(0))
    t = buf[:nctas].cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    names = ['start', 'setup', 'mma1st', 'mmaDone', 'epiWake', 'tmemLd1', 'epiDone', 'end']
    rel = rel[:, [0, 1, 4, 5, 6, 2, 3, 7]]
    print(name, 'ctas', nctas)
    for i, n in enumerate(names):
        col = rel[:, i]
        print('   %-8s min %7.2f  median %7.2f  max %7.2f us' % (n, col.min(), np.median(col), col.max()))
M, C = 457, 384
h = torch.randn(M, C, device=dev).bfloat16()
wq = (torch.randn(3*C, C, device=dev)/20).bfloat16()
qkv = torch.empty(M, 3*C, device=dev, dtype=torch.bfloat16)
trace('QKV bn64', lambda: ops.gemm(h, wq, in_dtype=ops.BF16, out=qkv, out_dtype=ops.BF16), 72)
ff = torch.randn(M, 4*C, device=dev).bfloat16(); w2 = (torch.randn(C, 4*C, device=dev)/40).bfloat16()
x = torch.randn(M, C, device=dev); b = torch.randn(C, device=dev)
trace('FF2 split auto', lambda: ops.gemm(ff, w2, bias=b, in_dtype=ops.BF16, out=x, out_dtype=ops.F32, accumulate=True), 144)
sys.exit(0)
# a weight-streaming shape: M=100 rows, N=147456, K=3072 (decoder conv.2)
h1 = torch.randn(100, 3072, device=dev).bfloat16(); w = (torch.randn(147456, 3072, device=dev)/55).bfloat16()
out = torch.empty(100, 147456, device=dev)
trace('conv2-like M=100 (first 1152 ctas)', lambda: ops.gemm(h1, w, in_dtype=ops.BF16, out=out, out_dtype=ops.F32, block_n=128), 1152)
