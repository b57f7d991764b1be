"""Per-CTA phase timeline (globaltimer stamps) of the Graphormer GEMMs of an all-architecture batch (M = 18 666)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from ghn3_b200 import ops, _lib as L
dev = 'cuda'
lib = L.load()
buf = torch.zeros(4096, 8, dtype=torch.int64, device=dev)


def trace(name, fn, nctas):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    buf.zero_()
    lib.ghn3_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.ghn3_debug_gemm_trace(ctypes.c_void_p(0))
    n = min(nctas, 4096)
    t = buf[:n].cpu().numpy().astype(np.float64)
    t = t[t[:, 0] > 0]
    names = ['start', 'setup', 'mma1st', 'mmaDone', 'epiWake', 'tmemLd1', 'epiDone', 'end']
    t = t[:, [0, 1, 4, 5, 6, 2, 3, 7]]
    d = np.diff(t, axis=1) / 1e3
    print('%s: %.1f us per launch, %d CTAs traced; median us per phase: %s; CTA lifetime median %.2f us' % (
        name, e0.elapsed_time(e1) * 100, len(t), ', '.join('%s->%s %.2f' % (names[i], names[i + 1], np.median(d[:, i]))
                                                            for i in range(7)),
        np.median((t[:, 7] - t[:, 0]) / 1e3)))


M, C = 18666, 384
h = torch.randn(M, C, device=dev).bfloat16()
wq = (torch.randn(3 * C, C, device=dev) / 20).bfloat16()
qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
trace('QKV  (9 x 146 tiles)', lambda: ops.gemm(h, wq, in_dtype=ops.BF16, out=qkv, out_dtype=ops.BF16, b_dynamic=False), 1314)
w1 = (torch.randn(4 * C, C, device=dev) / 20).bfloat16()
b1 = torch.randn(4 * C, device=dev)
ff = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
trace('FFN1 (12 x 146, bias + GELU)', lambda: ops.gemm(h, w1, bias=b1, act=ops.ACT_GELU, in_dtype=ops.BF16, out=ff, out_dtype=ops.BF16, b_dynamic=False), 1752)
w2 = (torch.randn(C, 4 * C, device=dev) / 40).bfloat16()
x = torch.randn(M, C, device=dev)
b2 = torch.randn(C, device=dev)
trace('FFN2 (3 x 146, K = 1536, residual)', lambda: ops.gemm(ff, w2, bias=b2, in_dtype=ops.BF16, out=x, out_dtype=ops.F32, accumulate=True, b_dynamic=False), 438)
wo = (torch.randn(C, C, device=dev) / 20).bfloat16()
trace('out-proj (3 x 146, residual)', lambda: ops.gemm(h, wo, bias=b2, in_dtype=ops.BF16, out=x, out_dtype=ops.F32, accumulate=True, b_dynamic=False), 438)
