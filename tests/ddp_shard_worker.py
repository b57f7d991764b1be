"""Worker of test_two_rank_sharded_optimizer (launched with torch.distributed.run, 2 ranks, NCCL):
(1) GradSync(shard=True): the in-place reduce-scatter leaves the MEAN over ranks in this rank's slice of each region;
(2) FusedAdamW.enable_sharding: three sharded steps (|g|^2 summed over ranks, AdamW on the own slices, all-gather of
    the parameters) against three replicated steps on the same gradients;
(3) Trainer end to end: sharded and replicated training of the same GHN on the same meta-batch give the same losses;
    gather_state() reassembles the moments."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghn3_b200 import GHN3, Graph, Trainer                      # noqa: E402
from ghn3_b200.optim import FusedAdamW                          # noqa: E402
from ghn3_b200.train import FLAT_PAD, GradSync                  # noqa: E402
from ghn3_b200.trainer import shard_meta_batch                  # noqa: E402
from ghn3_b200.weights import CONFIGS, procedural_state_dict    # noqa: E402
from tests import helpers as H                                  # noqa: E402

ARCHS = ['resnet18', 'squeezenet1_1']


def main():
    out = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = 'cuda'
    dist.init_process_group('nccl')
    cfg = CONFIGS['ghn3tiny']
    res = {}

    def make():
        ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='tf32')
        ghn.load_state_dict(procedural_state_dict(cfg, 0))
        return ghn.to(dev).train()

    # (1) reduce-scatter in place
    sync = GradSync(shard=True)
    assert sync.shard
    n = 5 * FLAT_PAD
    g = torch.Generator(device=dev).manual_seed(10 + rank)
    buf = torch.randn(n, device=dev, generator=g)
    ref = buf.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    sync.finish([sync.start(buf)], buf)
    a, b = sync.shard_of(0, n)
    res['rs_err'] = float((buf[a:b] - ref[a:b]).abs().max())

    # (2) sharded against replicated optimizer steps on identical gradients
    ga, gb = make(), make()
    oa = FusedAdamW(ga, lr=3e-3, weight_decay=0.05, max_grad_norm=0.5)
    ob = FusedAdamW(gb, lr=3e-3, weight_decay=0.05, max_grad_norm=0.5)
    assert oa.enable_sharding(sync)
    for step in range(3):
        gen = torch.Generator(device=dev).manual_seed(100 + step)            # same "averaged" gradient on every rank
        for pa, pb in zip(ga.parameters(), gb.parameters()):
            gr = torch.randn(pa.shape, device=dev, generator=gen)
            pa.grad, pb.grad = gr, gr.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    res['opt_err'] = max(H.max_rel_err(pa, pb) for pa, pb in zip(ga.parameters(), gb.parameters()))
    oa.gather_state()
    res['moment_err'] = max(float((oa.exp_avg - ob.exp_avg).abs().max()),
                            float((oa.exp_avg_sq - ob.exp_avg_sq).abs().max()))
    # the sharded GHN still predicts (device weight copies were rebuilt over the flat parameter buffer)
    with torch.no_grad():
        ga.eval()
        gb.eval()
        ma = ga(H.build_model('resnet18').to(dev), Graph.from_record(H.graph_records()['resnet18']))
        mb = gb(H.build_model('resnet18').to(dev), Graph.from_record(H.graph_records()['resnet18']))
        torch.cuda.synchronize()
        res['predict_err'] = max(H.max_rel_err(x, y) for x, y in zip(ma.parameters(), mb.parameters()))
    del ga, gb, oa, ob

    # (3) Trainer: sharded and replicated runs of the same 2-graph meta-batch
    mine = [ARCHS[i] for i in shard_meta_batch(len(ARCHS), rank, world)]
    losses = {}
    for mode in (True, False):
        ghn = make()
        trainer = Trainer(ghn, opt='adamw', opt_args={'lr': 1e-3, 'weight_decay': 1e-2, 'shard_optimizer': mode},
                          grad_clip=5, device=dev)
        assert (getattr(trainer._optimizer, '_sync', None) is not None) == mode
        nets = [H.build_model(a).to(dev) for a in mine]
        graphs = [Graph.from_record(H.graph_records()[a]) for a in mine]

        def loss_fn(models):
            loss = 0
            for a, net in zip(mine, models):
                gen = torch.Generator().manual_seed(ARCHS.index(a))
                for p in net.parameters():
                    loss = loss + (p * torch.randn(p.shape, generator=gen).to(dev)).sum()
            return loss
        ls = []
        for _ in range(4):
            trainer.update(None, None, graphs=graphs, models=nets, loss_fn=loss_fn)
            ls.append(trainer.metrics['loss'].avg)
            trainer.reset_metrics()
        losses[mode] = ls
        if mode:
            chk = torch.stack([p.detach().double().sum() for p in ghn.parameters()]).sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            res['params_equal'] = bool((lo == hi).item())
    res['losses_sharded'], res['losses_replicated'] = losses[True], losses[False]
    if rank == 0:
        torch.save(res, os.path.join(out, 'res.pt'))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
