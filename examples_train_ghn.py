"""
GHN-3 training on synthetic data -- the loop of the reference's train_ghn_ddp.py (lines 84-160) on the B200 path:

    python examples_train_ghn.py --cfg ghn3tm8 --steps 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        examples_train_ghn.py --cfg ghn3xlm16 --meta-batch 8

Every step draws a meta-batch of DeepNets-1M-style architectures (ghn3_b200.deepnets.NetGenerator; same seed on every
rank, each rank keeps its contiguous share as in train_ghn_ddp.py:92), predicts their parameters with keep_grads=True,
runs them on a batch of synthetic images, and back-propagates the mean cross-entropy into the GHN (hand-written
adjoint, gradients averaged over ranks inside it, clip + AdamW in one kernel). There is no ImageNet / DeepNets-1M
download: images are random, so the loss only demonstrates that the step optimises.
"""
import argparse
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cfg', default='ghn3tm8', help='ghn3tm8 | ghn3sm8 | ghn3lm8 | ghn3xlm16 | ghn3tiny')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'tf32'])
    ap.add_argument('--meta-batch', type=int, default=8)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--images', type=int, default=32, help='images per rank and step')
    ap.add_argument('--image-size', type=int, default=64)
    ap.add_argument('--lr', type=float, default=4e-4)
    ap.add_argument('--wd', type=float, default=1e-2)
    ap.add_argument('--pool', type=int, default=4, help='number of distinct meta-batches cycled through')
    args = ap.parse_args()

    import torch.distributed as dist
    from ghn3_b200 import GHN3, GraphBatch, Trainer
    from ghn3_b200.deepnets import NetGenerator
    from ghn3_b200.trainer import shard_meta_batch
    from ghn3_b200.weights import CONFIGS, procedural_state_dict

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dev = torch.device('cuda', torch.cuda.current_device())
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    cfg = CONFIGS[args.cfg]
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=args.dtype)
    ghn.load_state_dict(procedural_state_dict(cfg, 0))          # same initial weights on every rank
    trainer = Trainer(ghn, opt='adamw', opt_args={'lr': args.lr, 'weight_decay': args.wd}, grad_clip=5,
                      predparam_wd=3e-5, amp=True, device=dev, log_interval=5)

    gen = NetGenerator(seed=0, max_params=5e6)
    pool = []
    for _ in range(args.pool):                                   # host tracing happens once per architecture
        pairs = gen.sample(args.meta_batch, input_size=args.image_size)
        mine = [pairs[i] for i in shard_meta_batch(args.meta_batch, rank, world)]
        pool.append((GraphBatch([g for _, g in mine], dense=True).to_device(dev), [n.to(dev) for n, _ in mine]))

    g = torch.Generator().manual_seed(1 + rank)
    t0 = time.time()
    for step in range(args.steps):
        graphs, nets = pool[step % len(pool)]
        images = torch.randn(args.images, 3, args.image_size, args.image_size, generator=g)
        targets = torch.randint(0, cfg['num_classes'], (args.images,), generator=g)
        metrics = trainer.update(images, targets, graphs=graphs, models=nets)
        if rank == 0 and (step + 1) % 5 == 0:
            torch.cuda.synchronize()
            print('step %3d  loss %.4f  top5 %.2f  %.1f graphs/s' % (
                step + 1, metrics['loss'].avg, metrics['top5'].avg,
                args.meta_batch * (step + 1) / (time.time() - t0)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
