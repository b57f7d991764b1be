"""Host-side cost of one Trainer.update on the bench's training meta-batch (cProfile, no device synchronisation)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import GHN3, GraphBatch, Trainer
from ghn3_b200.deepnets import NetGenerator
from ghn3_b200.weights import CONFIGS, procedural_state_dict

dev = torch.device('cuda')
cfg = CONFIGS['ghn3xlm16']
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
ghn.load_state_dict(procedural_state_dict(cfg, 0))
trainer = Trainer(ghn, opt='adamw', opt_args={'lr': 4e-4, 'weight_decay': 1e-2}, grad_clip=5, device=dev)
pairs = NetGenerator(seed=0).sample(bench.TRAIN_META_BATCH)
graphs = GraphBatch([g for _, g in pairs], dense=True).to_device(dev)
nets = [n.to(dev) for n, _ in pairs]
loss_fn = lambda models: ghn.last_program.pred_flat.sum() * 1e-3
for _ in range(3):
    trainer.update(None, None, graphs=graphs, models=nets, loss_fn=loss_fn)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    trainer.update(None, None, graphs=graphs, models=nets, loss_fn=loss_fn)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host enqueue time per step %.2f ms; wall per step incl. drain %.2f ms' % ((t1 - t0) * 100, (t2 - t0) * 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    trainer.update(None, None, graphs=graphs, models=nets, loss_fn=loss_fn)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(16)
