"""
Config 3 of BASELINE.json: predict the parameters of ALL torchvision classification architectures with one GHN-3,
architectures sharded across the GPUs of one box (LPT on a byte/FLOP cost, no data-path collective).

    python examples_predict_all.py --ghn ghn3lm8                         # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples_predict_all.py --ghn ghn3lm8

Weights are procedural random-init (no network access); graphs come from the committed fixture records unless
--trace is given (then every model is traced on the host by ghn3_b200.tracer).
"""
import argparse
import gzip
import json
import os
import time

import torch
import torch.distributed as dist
import torchvision.models as tvm

from ghn3_b200 import GHN3, Graph
from ghn3_b200.shard import architecture_cost, shard_lpt
from ghn3_b200.weights import CONFIGS, procedural_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ghn', default='ghn3lm8', choices=sorted(CONFIGS))
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'tf32', 'tf32x1'])
    ap.add_argument('--trace', action='store_true')
    ap.add_argument('--repeat', type=int, default=3)
    ap.add_argument('--batch', type=int, default=1, help='architectures per ghn(...) call (varlen-packed batch)')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    with gzip.open(os.path.join(HERE, 'tests', 'golden', 'graphs_tv.json.gz'), 'rt') as f:
        records = json.load(f)
    names = sorted(records)
    costs = [architecture_cost(records[n]['n'], records[n]['n_params']) for n in names]
    mine = shard_lpt(costs, world)[rank]
    cfg = CONFIGS[args.ghn]
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=args.dtype)
    ghn.load_state_dict(procedural_state_dict(cfg, 0))
    ghn = ghn.to(dev).eval()
    models, graphs = [], []
    for i in mine:
        kw = {'init_weights': False} if names[i] in ('googlenet', 'inception_v3') else {}
        m = getattr(tvm, names[i])(**kw)
        if names[i] == 'inception_v3':
            m.expected_input_sz = 299
        graphs.append(Graph(m, verbose=False) if args.trace else Graph.from_record(records[names[i]]))
        models.append(m.to(dev))
    best = None
    with torch.no_grad():
        for _ in range(args.repeat):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(0, len(models), args.batch):
                ghn(models[i:i + args.batch], graphs[i:i + args.batch])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = dt if best is None else min(best, dt)
    if rank == 0:
        print(json.dumps({'ghn': args.ghn, 'dtype': args.dtype, 'n_gpus': world, 'architectures': len(names),
                          'seconds': best, 'models_per_s': len(names) / best}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
