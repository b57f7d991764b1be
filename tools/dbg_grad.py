import sys; sys.path.insert(0, '/root/repo')
import torch, numpy as np
import tests.test_train_gpu as T
from tests.test_train_gpu import CONFIGS, H, procedural_state_dict
_oracle_grads_named, _cuda_grads = T._oracle_grads_named, T._cuda_grads
archs = ['resnet18']
cfg = CONFIGS['ghn3tiny']
recs = [H.graph_records()[a] for a in archs]
sd_grads, all_R, ref_loss = _oracle_grads_named(cfg, archs, recs)
gb, lb = _cuda_grads('ghn3tiny', 'bf16', archs, recs, all_R)
gt, lt = _cuda_grads('ghn3tiny', 'tf32', archs, recs, all_R)
print('loss', ref_loss, lb, lt)
for k in sd_grads:
    r = sd_grads[k]
    if r is None or float(r.abs().max()) == 0: continue
    e_b = float((gb[k].float().cpu() - r).norm() / r.norm())
    e_t = float((gt[k].float().cpu() - r).norm() / r.norm())
    e_bt = float((gb[k].float().cpu() - gt[k].float().cpu()).norm() / r.norm())
    print('%-45s bf16-vs-fp32 %.4f  tf32-vs-fp32 %.2e  bf16-vs-tf32 %.4f' % (k, e_b, e_t, e_bt))
