"""One Trainer step of bench.py's training leg (same meta-batch, keyed stub loss, predparam_wd, fused clip + AdamW)
delimited by cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists.
  python tools/train_step_profile.py [--graphs 8] [--dtype bf16]     (--graphs 1 = the per-rank share at 8 GPUs)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import GHN3, GraphBatch, Trainer
from ghn3_b200.deepnets import NetGenerator
from ghn3_b200.weights import CONFIGS, procedural_state_dict

ap = argparse.ArgumentParser()
ap.add_argument('--graphs', type=int, default=bench.TRAIN_META_BATCH)
ap.add_argument('--dtype', default='bf16')
ap.add_argument('--steps', type=int, default=1)
a = ap.parse_args()
dev = torch.device('cuda')
cfg = CONFIGS['ghn3xlm16']
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=a.dtype)
ghn.load_state_dict(procedural_state_dict(cfg, 0))
ghn = ghn.to(dev).train()
pairs = NetGenerator(seed=0).sample(bench.TRAIN_META_BATCH)[:a.graphs]
ids = list(range(a.graphs))
trainer = Trainer(ghn, opt='adamw', opt_args={'lr': 4e-4, 'weight_decay': 1e-2}, grad_clip=5, device=dev,
                  predparam_wd=3e-5)
graphs = GraphBatch([g for _, g in pairs], dense=True).to_device(dev)
nets = [n.to(dev) for n, _ in pairs]
state = {}


def loss_fn(models):
    flat = ghn.last_program.pred_flat
    if 'R' not in state:
        state['R'] = bench._keyed_loss_weights(models, ids, flat)
    return (flat * state['R']).sum()


for _ in range(3):
    trainer.update(None, None, graphs=graphs, models=nets, loss_fn=loss_fn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
for _ in range(a.steps):
    trainer.update(None, None, graphs=graphs, models=nets, loss_fn=loss_fn)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('%d graphs: %.3f ms per step' % (a.graphs, e0.elapsed_time(e1) / a.steps))
