"""Per-CTA phase timeline of an fc-like grouped launch (non-persistent kernel, GHN3_NO_PERSISTENT=1)."""
import os, sys, ctypes
os.environ['GHN3_NO_PERSISTENT'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from ghn3_b200 import ops, _lib as L
dev = 'cuda'
lib = L.load()
C = 384
a = torch.randn(457, C, device=dev).bfloat16()
w = (torch.randn(256 * 4 * C, C, device=dev) / 20).bfloat16()
bias = torch.randn(256 * 4 * C, device=dev)
probs = np.zeros(256, dtype=[('a_row0', 'i4'), ('b_row0', 'i4'), ('m', 'i4'), ('n', 'i4'), ('d_off', 'i8'), ('ldd', 'i4'), ('bias_off', 'i4')])
tiles = []
off = 0
for p in range(256):
    probs[p] = (0, p * 4 * C, 38, 4 * C, off, 4 * C, p * 4 * C)
    off += 38 * 4 * C
    for nt in range(12):
        tiles.append((p, 0, nt, 0))
out = torch.empty(off, device=dev, dtype=torch.bfloat16)
pd = torch.from_numpy(probs.view(np.uint8).copy()).to(dev)
td = torch.tensor(tiles, dtype=torch.int32, device=dev)
fn = lambda: ops.gemm(a, w, bias=bias, act=ops.ACT_RELU, in_dtype=ops.BF16, out=out, out_dtype=ops.BF16, problems=pd, tiles=td, b_dynamic=False)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(5):
    junk.fill_(1); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print('fc-like non-persistent: %.1f us  (%.0f GB/s)' % (np.median(ts) * 1e3, w.numel() * 2 / np.median(ts) / 1e6))
buf = torch.zeros(4096, 8, dtype=torch.int64, device=dev)
junk.fill_(1); torch.cuda.synchronize()
lib.ghn3_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr())); fn(); torch.cuda.synchronize(); lib.ghn3_debug_gemm_trace(ctypes.c_void_p(0))
t = buf[:3072].cpu().numpy().astype(np.float64)
names = ['start', 'setup', 'prepDone', 'epiDone', 'mma1st', 'mmaDone', 'epiWake', 'end']
d = {n: t[:, i] for i, n in enumerate(names)}
t0 = d['start'].min()
print('kernel span %.1f us' % ((d['end'].max() - t0) / 1e3))
for a_, b_ in [('start', 'setup'), ('setup', 'prepDone'), ('setup', 'mma1st'), ('mma1st', 'mmaDone'), ('mmaDone', 'epiWake'), ('epiWake', 'epiDone'), ('epiDone', 'end'), ('start', 'end')]:
    dd = (d[b_] - d[a_]) / 1e3
    print('  %-8s -> %-8s median %6.2f  p90 %6.2f  max %6.2f us' % (a_, b_, np.median(dd), np.percentile(dd, 90), dd.max()))
os.environ.pop('GHN3_NO_PERSISTENT')
