"""Bring-up diagnostics for the tcgen05 GEMM (run on the GPU box): structured inputs make layout bugs visible."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ghn3_b200 import ops

def report(name, out, ref):
    err = (out.double() - ref.double()).abs()
    print('%-28s max_abs_err=%.3e ref_max=%.3e bad=%d/%d' % (name, err.max().item(), ref.abs().max().item(),
                                                              int((err > 1e-3 * ref.abs().max()).sum()), err.numel()))
    if err.max() > 1e-3 * ref.abs().max():
        bad = (err > 1e-3 * ref.abs().max()).nonzero()
        print('   first bad idx', bad[:8].tolist())
        print('   bad rows', sorted(set(bad[:, 0].tolist()))[:20], 'bad cols', sorted(set(bad[:, 1].tolist()))[:20])
        print('   out[0,:8]', out[0, :8].tolist()); print('   ref[0,:8]', ref[0, :8].tolist())

for dtype, name in ((ops.BF16, 'bf16'), (ops.TF32, 'tf32')):
    for (m, n, k) in [(128, 128, 64), (128, 128, 32), (128, 128, 256), (64, 32, 64), (256, 256, 512)]:
        torch.manual_seed(0)
        a = torch.randint(-2, 3, (m, k), device='cuda').float()
        b = torch.randint(-2, 3, (n, k), device='cuda').float()
        ai, bi = (a.bfloat16(), b.bfloat16()) if dtype == ops.BF16 else (a, b)
        try:
            out = ops.gemm(ai, bi, in_dtype=dtype, out_dtype=ops.F32)
            torch.cuda.synchronize()
            report('%s int %dx%dx%d' % (name, m, n, k), out, a @ b.t())
        except Exception as e:
            print(name, (m, n, k), 'EXC', e)
            sys.exit(1)
print('done')
