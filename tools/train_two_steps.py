"""Two training steps (the second one delimited by cudaProfilerStart/Stop) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ghn3_b200 import GHN3, Graph, GraphBatch
from ghn3_b200.weights import CONFIGS, procedural_state_dict

dtype = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
dev = torch.device('cuda')
cfg = CONFIGS['ghn3xlm16']
records = bench.load_records()
ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=dtype)
ghn.load_state_dict(procedural_state_dict(cfg, 0))
ghn = ghn.to(dev).train()
from ghn3_b200.deepnets import NetGenerator
from ghn3_b200.optim import FusedAdamW
pairs = NetGenerator(seed=0).sample(bench.TRAIN_META_BATCH)          # the bench's training meta-batch
graphs = GraphBatch([g for _, g in pairs], dense=True).to_device(dev)
nets = [n.to(dev) for n, _ in pairs]
opt = FusedAdamW(ghn, lr=4e-4, weight_decay=1e-2, max_grad_norm=5)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    opt.zero_grad(set_to_none=True)
    ghn(nets, graphs, keep_grads=True, reduce_graph=True)
    (ghn.last_program.pred_flat.sum() * 1e-3).backward()
    opt.step()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
