"""Worker of test_two_rank_gradients_equal_single_rank_mean (launched with torch.distributed.run, 2 ranks, NCCL)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghn3_b200 import GHN3, Graph                               # noqa: E402
from ghn3_b200.train import enable_grad_sync                    # noqa: E402
from ghn3_b200.trainer import shard_meta_batch                  # noqa: E402
from ghn3_b200.weights import CONFIGS, procedural_state_dict    # noqa: E402
from tests import helpers as H                                  # noqa: E402

ARCHS = ['resnet18', 'squeezenet1_1']


def grads(ghn, archs, dev):
    nets = [H.build_model(a).to(dev) for a in archs]
    graphs = [Graph.from_record(H.graph_records()[a]) for a in archs]
    out = ghn(nets, graphs, keep_grads=True)
    loss = 0
    for a, net in zip(archs, out):
        g = torch.Generator().manual_seed(ARCHS.index(a))
        for p in net.parameters():
            loss = loss + (p * torch.randn(p.shape, generator=g).to(dev)).sum()
    (loss / len(archs)).backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().cpu() for k, p in ghn.named_parameters()}


def main():
    out = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = 'cuda'
    dist.init_process_group('nccl')
    cfg = CONFIGS['ghn3tiny']

    def make():
        ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='tf32')
        ghn.load_state_dict(procedural_state_dict(cfg, 0))
        return ghn.to(dev).train()
    ghn = enable_grad_sync(make())
    mine = [ARCHS[i] for i in shard_meta_batch(len(ARCHS), rank, world)]
    g = grads(ghn, mine, dev)
    if rank == 0:
        torch.save(g, os.path.join(out, 'rank0.pt'))
        torch.save(grads(make(), ARCHS, dev), os.path.join(out, 'single.pt'))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
