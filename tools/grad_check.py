"""Per-tensor gradient agreement (relative L2, cosine, max-rel) of the CUDA training path against oracle autograd."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import test_train_gpu as T
from tests import helpers as H
from ghn3_b200.weights import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument('--cfg', default='ghn3tiny')
ap.add_argument('--archs', default='resnet18')
ap.add_argument('--dtype', default='bf16')
a = ap.parse_args()
archs = a.archs.split(',')
cfg = CONFIGS[a.cfg]
recs = [H.graph_records()[x] for x in archs]
ref, all_R, ref_loss = T._oracle_grads_named(cfg, archs, recs)
got, loss = T._cuda_grads(a.cfg, a.dtype, archs, recs, all_R)
print('loss', loss, ref_loss)
rows = []
for k, r in ref.items():
    g = got[k].float().cpu()
    l2 = float((g - r).norm() / (r.norm() + 1e-30))
    cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
    mx = float((g - r).abs().max() / (r.abs().max() + 1e-30))
    rows.append((l2, cos, mx, k))
for l2, cos, mx, k in sorted(rows, reverse=True)[:60]:
    print('%-45s relL2 %.4f cos %.5f maxrel %.4f' % (k, l2, cos, mx))
