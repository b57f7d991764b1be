"""Where does the bf16 gradient error enter? conv.2 dgrad on the program's own buffers vs torch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tests.test_train_gpu as T
from tests.test_train_gpu import CONFIGS, H, procedural_state_dict
from ghn3_b200 import GHN3, Graph
DEV = 'cuda'
cfg = CONFIGS['ghn3tiny']
rec = H.graph_records()['resnet18']
res = {}
for dtype in ('bf16', 'tf32'):
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=dtype)
    ghn.load_state_dict(procedural_state_dict(cfg, 0))
    ghn = ghn.to(DEV).train()
    model = ghn(H.build_model('resnet18').to(DEV), Graph.from_record(rec), keep_grads=True)
    torch.manual_seed(3)
    loss = sum((p * torch.randn_like(p)).sum() for p in model.parameters())
    loss.backward()
    torch.cuda.synchronize()
    prog = ghn.last_program
    b = prog.bwd
    res[dtype] = dict(X=b.X.float().clone(), dh1=b.dh1.float().clone(), dh0=b.dh0.float().clone(), h1=prog.h1.float().clone(),
                      dwout=b.dwout.clone(), W2=ghn.decoder.conv[2].weight.detach().clone(),
                      c2T=b.wt['c2_wT'].float().clone())
rb, rt = res['bf16'], res['tf32']
rel = lambda a, b_: float((a - b_).norm() / b_.norm())
print('dwout  bf16 vs tf32:', rel(rb['dwout'], rt['dwout']))
print('X      bf16 vs tf32:', rel(rb['X'], rt['X']))
print('dh1    bf16 vs tf32:', rel(rb['dh1'], rt['dh1']))
print('dh0    bf16 vs tf32:', rel(rb['dh0'], rt['dh0']))
# recompute dh1 from the bf16 program's own X and W2 in fp32
W2 = rb['W2']
ref = (rb['X'] @ W2.bfloat16().float()) * (rb['h1'] > 0)
print('dh1 bf16 kernel vs torch on the same bf16 operands:', rel(rb['dh1'], ref))
ref32 = (rt['X'] @ W2) * (rt['h1'] > 0)
print('dh1 tf32 kernel vs torch fp32:', rel(rt['dh1'], ref32))
print('torch: bf16-operand dh1 vs fp32-operand dh1:', rel(ref, ref32))
mask_diff = float(((rb['h1'] > 0) != (rt['h1'] > 0)).float().mean())
print('relu mask mismatch fraction (h1):', mask_diff)
print('c2T (bf16 transposed copy) vs W2^T:', rel(rb['c2T'][:, :W2.shape[0]], W2.t()))
X32, Xb, Wb = rt['X'], rb['X'], W2.bfloat16().float()
full = X32 @ W2
print('||X (Wb - W)|| / ||X W|| =', rel(X32 @ Wb, full), '   ||(Xb - X) W|| / ||X W|| =', rel(Xb @ W2, full))
rn = full.norm(dim=1); xn = X32.norm(dim=1)
amp = (xn * W2.norm() / (W2.shape[0] ** 0.5)) / rn.clamp_min(1e-30)
print('rows', X32.shape[0], ' cancellation factor |x| |W|_F/sqrt(K) / |x W| : median %.1f  max %.1f' % (float(amp.median()), float(amp.max())))
nz = (X32 != 0).float().sum(1)
print('nonzeros per row: min %d median %d max %d of %d' % (int(nz.min()), int(nz.median()), int(nz.max()), X32.shape[1]))
e_row = ((Xb @ Wb) - full).norm(dim=1) / rn.clamp_min(1e-30)
big = rn > rn.max() * 0.05
print('per-row relative error: median %.4f, rows carrying the norm: %d, their error median %.4f' % (float(e_row.median()), int(big.sum()), float(e_row[big].median())))
i = int(rn.argmax()); x = X32[i]; print('largest row: |x|=%.3e nz=%d mean=%.3e std=%.3e  |xW|=%.3e' % (float(x.norm()), int(nz[i]), float(x[x != 0].mean()), float(x[x != 0].std()), float(rn[i])))
