"""
One GHN training step on the B200 path -- the GHN branch of the reference's `Trainer.update`
(ghn3/trainer.py:238-411) with the same argument meaning:

    trainer = Trainer(ghn, opt='adamw', opt_args={'lr': 4e-4, 'weight_decay': 1e-2}, grad_clip=5, predparam_wd=3e-5)
    metrics = trainer.update(images, targets, graphs=graph_batch)       # graph_batch.nets = the target networks

What differs from the reference, by design:
  * no autocast / GradScaler: the GHN computes in bf16 or error-compensated tf32 with fp32 accumulation and fp32
    gradients (`GHN3(compute_dtype=...)`), so the fp16 range workarounds (trainer.py:343-379) have nothing to do;
  * data parallelism is not DistributedDataParallel: the hand-written backward pass all-reduces its one flat
    gradient buffer itself (ghn3_b200.train.GradSync), overlapping the decoder gradients with the Graphormer adjoint;
  * metrics are kept on the device and averaged across ranks with ONE packed all-reduce per step instead of the
    4-5 `.item()` + all_gather round trips (trainer.py:381-388, ddp_utils.py:84-93).
"""
import os

import torch
import torch.nn as nn

from .train import enable_grad_sync


def is_ddp():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_meta_batch(n_items, rank, world_size):
    """Indices of the meta-batch items owned by `rank`: contiguous equal shares (reference train_ghn_ddp.py:92 gives
    every rank meta_batch_size // world_size graphs)."""
    if n_items % world_size != 0:
        raise ValueError('meta batch size %d must be divisible by the number of ranks %d (train_ghn_ddp.py:92)'
                         % (n_items, world_size))
    per = n_items // world_size
    return list(range(rank * per, (rank + 1) * per))


class AvgMeter:
    """Running average whose updates may be DEVICE scalars: they are accumulated on the device and only read back
    (one synchronisation) when `.avg` / `.sum` is asked for -- a training step never waits for its own metrics."""

    def __init__(self):
        self._sum, self.cnt = 0.0, 0
        self._pending = None

    def update(self, val, n=1):
        if torch.is_tensor(val):
            v = val.detach().double() * n
            self._pending = v if self._pending is None else self._pending + v
        else:
            self._sum += float(val) * n
        self.cnt += n

    def _flush(self):
        if self._pending is not None:
            self._sum += float(self._pending)
            self._pending = None

    @property
    def sum(self):
        self._flush()
        return self._sum

    @sum.setter
    def sum(self, v):
        self._pending = None
        self._sum = v

    @property
    def avg(self):
        return self.sum / max(self.cnt, 1)


class Trainer:
    def __init__(self, model, opt='adamw', opt_args=None, scheduler=None, n_batches=None, grad_clip=5,
                 device='cuda', log_interval=100, label_smoothing=0, predparam_wd=0, verbose=False, amp=False,
                 **unused):
        self.criterion = nn.CrossEntropyLoss(label_smoothing=label_smoothing)
        self.n_batches, self.grad_clip, self.device = n_batches, grad_clip, device
        self.log_interval, self.predparam_wd, self.verbose = log_interval, predparam_wd, verbose
        # amp: the TARGET networks run under bf16 autocast (the reference uses fp16 autocast + GradScaler,
        # trainer.py:269,343-379; bf16 needs no loss scaling). The GHN itself always computes in its compute_dtype.
        self.amp = amp
        self.ddp = is_ddp()
        model.to(device)
        self._model = model
        model.direct_grads = True          # .grad = views of the backward pass's flat buffer (see train._PredictFn)
        opt_args = dict(opt_args or {})
        # shard_optimizer (default on under NCCL data parallelism with the fused AdamW): the gradient is reduce-scattered
        # instead of all-reduced, every rank updates 1/N of the parameters and the result is all-gathered
        shard = opt_args.pop('shard_optimizer', os.environ.get('GHN3_SHARD_OPTIMIZER', 'auto'))
        if isinstance(shard, str):
            if shard.lower() == 'auto':              # pays from 4 ranks on (measured: 2 ranks -2 %)
                import torch.distributed as dist
                shard = self.ddp and dist.get_world_size() >= 4
            else:
                shard = shard.lower() not in ('0', 'false', 'off', 'no')
        shard = bool(shard)
        if self.ddp:
            enable_grad_sync(model, shard=shard and isinstance(opt, str) and opt.lower() == 'adamw'
                             and opt_args.get('fused_clip', True) and hasattr(model, 'decoder_1d'))
        if self.ddp:
            # DistributedDataParallel broadcasts rank 0's parameters and buffers at construction (reference
            # trainer.py:136); ranks built from different seeds would otherwise diverge silently
            import torch.distributed as dist
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, 0)
            if hasattr(model, 'weights_updated'):
                model.weights_updated()
        params = [p for p in model.parameters() if p.requires_grad]
        self._fused_clip = False
        if isinstance(opt, str):
            name = opt.lower()
            fused_clip = opt_args.pop('fused_clip', True)
            if name != 'sgd':
                opt_args.pop('momentum', None)              # as the reference does for adam / adamw (trainer.py:176)
            if name == 'adamw' and fused_clip and hasattr(model, 'decoder_1d'):
                # clipping + AdamW in one pass over the backward pass's flat gradient buffer (ghn3_adamw)
                from .optim import FusedAdamW
                self._optimizer = FusedAdamW(model, max_grad_norm=grad_clip, **opt_args)
                self._fused_clip = True
                if self.ddp:
                    self._optimizer.enable_sharding(getattr(model, '_grad_sync', None))
            elif name == 'sgd':
                opt_args.setdefault('momentum', 0.9)
                self._optimizer = torch.optim.SGD(params, **opt_args)
            elif name in ('adam', 'adamw'):
                opt_args.setdefault('fused', True)          # one multi-tensor kernel per chunk of the 291 tensors
                self._optimizer = (torch.optim.Adam if name == 'adam' else torch.optim.AdamW)(params, **opt_args)
            else:
                raise NotImplementedError(opt)
        else:
            self._optimizer = opt
        self._scheduler = self._make_scheduler(scheduler, unused.get('scheduler_args'), unused.get('epochs', 300),
                                               opt_args.get('lr'))
        self._step = 0
        self.skipped_updates = 0
        self.reset_metrics()

    def save(self, path, epoch=0, step=0, config=None):
        """Checkpoint in the reference's layout (trainer.py:413-426): {'state_dict', 'optimizer', 'epoch', 'step',
        **config}; `from_pretrained(path)` loads it back. Rank 0 only under torch.distributed."""
        if self.ddp:
            import torch.distributed as dist
            if hasattr(self._optimizer, 'gather_state'):
                self._optimizer.gather_state()          # collective: every rank, before rank 0 writes
            if dist.get_rank() != 0:
                return
        self._model.flush()
        self.check_finite()                    # never checkpoint a state whose last steps were non-finite
        sd = {'state_dict': self._model.state_dict(), 'optimizer': self._optimizer.state_dict(), 'epoch': epoch,
              'step': step}
        sd.update(config or {'config': {k: self._model.config[k] for k in ('max_shape', 'num_classes', 'hid', 'heads',
                                                                          'layers', 'layernorm')} |
                             {'weight_norm': self._model.weight_norm, 've': self._model.ve}})
        torch.save(sd, path)

    def reset_metrics(self, epoch=0):
        self._step = 0
        self.metrics = {'loss': AvgMeter(), 'top1': AvgMeter(), 'top5': AvgMeter()}
        if self.predparam_wd > 0:
            self.metrics['loss_predwd'] = AvgMeter()

    def scheduler_step(self):
        if self._scheduler is not None:
            self._scheduler.step()

    def _make_scheduler(self, scheduler, scheduler_args, epochs, lr):
        """Scheduler objects pass through; the reference's strings (trainer.py:180-207: 'cosine-warmup[-steps5-init_lr1e-5]',
        'cosine', 'step', 'mstep') are built here over the (fused) optimizer."""
        if scheduler is None or not isinstance(scheduler, str):
            return scheduler
        import math
        from torch.optim.lr_scheduler import CosineAnnealingLR, LambdaLR, MultiStepLR, StepLR
        if scheduler.startswith('cosine-warmup'):
            def parse_arg(arg, default):
                p = scheduler.find(arg)
                if p <= 0:
                    return default
                p_end = scheduler[p:].find('-')
                return float(scheduler[p + len(arg):len(scheduler) if p_end == -1 else p + p_end])
            warmup_steps = int(parse_arg('steps', 5))
            warmup_lr = parse_arg('init_lr', 1e-5) / lr

            def lr_lambda(step):
                if step < warmup_steps - 1:
                    return warmup_lr + (1 - warmup_lr) * step / max(warmup_steps - 1, 1)
                progress = float(step - warmup_steps) / float(max(1, epochs - warmup_steps))
                return max(0.0, 0.5 * (1. + math.cos(math.pi * progress)))
            return LambdaLR(self._optimizer, lr_lambda=lr_lambda)
        if scheduler == 'cosine':
            return CosineAnnealingLR(self._optimizer, epochs)
        if scheduler == 'step':
            return StepLR(self._optimizer, **(scheduler_args or {}))
        if scheduler == 'mstep':
            return MultiStepLR(self._optimizer, **(scheduler_args or {}))
        raise NotImplementedError('scheduler %r (the reference knows cosine-warmup*, cosine, step, mstep)' % scheduler)

    # ------------------------------------------------------------------------------------------------------------
    def update(self, images, targets, graphs=None, models=None, loss_fn=None):
        """
        One step (reference trainer.py:238-411). `graphs.nets` (or `models`) are the target networks of this rank's
        share of the meta-batch. `loss_fn(models) -> scalar` replaces the image forward (used by the GHN-only
        benchmark); by default every network runs on `images` and the cross-entropy losses are averaged.
        Returns self.metrics.
        """
        ghn = self._model
        if not ghn.training:
            ghn.train()
        self._optimizer.zero_grad(set_to_none=True)
        if models is None:
            models = list(getattr(graphs, 'nets', None) or [])
        if not models:
            raise ValueError('Trainer.update needs the target networks (graphs.nets or models=...)')
        models = ghn(models, graphs, bn_track_running_stats=True, keep_grads=True, reduce_graph=True)
        models = models if isinstance(models, (list, tuple)) else [models]
        loss_predwd = None
        if self.predparam_wd > 0:
            # predparam_wd * sum_p ||p||_F (trainer.py:288-294) over the predicted tensors, on the device: three kernel
            # passes over the flat prediction buffer instead of one torch.norm (+ backward) per tensor
            from .train import predicted_param_decay
            loss_predwd = predicted_param_decay(ghn, self.predparam_wd)
        logits = None
        if loss_fn is not None:
            loss = loss_fn(models)
        else:
            targets = targets.to(self.device, non_blocking=True)
            images = images.to(self.device, non_blocking=True)
            loss, logits = 0, []
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.amp):
                for model in models:
                    out = model(images)
                    y = out[0] if isinstance(out, tuple) else out
                    loss = loss + self.criterion(y.float(), targets)
                    logits.append(y.detach().float())
            logits = torch.stack(logits)
        if loss_predwd is not None:
            loss = loss + loss_predwd
        loss = loss / len(models)                                   # mean over this rank's models (trainer.py:327)
        loss.backward()                                             # GHN adjoint + gradient all-reduce inside
        if self.grad_clip > 0 and not self._fused_clip:
            nn.utils.clip_grad_norm_([p for g in self._optimizer.param_groups for p in g['params']], self.grad_clip)
        if self._fused_clip and not self.ddp:
            # non-finite guard on the device (reference trainer.py:240-257 checks the loss before the update): the
            # fused step skips parameters AND moments when |g|^2 or the loss is not finite. Under data parallelism only
            # |g|^2 of the averaged gradient is used, so that every rank takes the same decision.
            self._optimizer.loss_for_guard = loss
        self._optimizer.step()

        # metrics: one packed device tensor, one all-reduce, one host read
        vals = [loss.detach().float()]
        if loss_predwd is not None:
            vals.append(loss_predwd.detach().float() if torch.is_tensor(loss_predwd) else torch.tensor(0.0))
        if logits is not None:
            tg = targets.view(1, -1).expand(len(logits), -1).reshape(-1)
            lg = logits.reshape(-1, logits.shape[-1])
            top = lg.topk(min(5, lg.shape[-1]), 1).indices
            hit = top.eq(tg.view(-1, 1))
            vals += [hit[:, :1].any(1).float().mean() * 100, hit.any(1).float().mean() * 100]
        packed = torch.stack([v.to(self.device) for v in vals])
        if self.ddp:
            import torch.distributed as dist
            dist.all_reduce(packed)
            packed = packed / dist.get_world_size()
        n = 1 if logits is None else logits.shape[0] * logits.shape[1]
        self.metrics['loss'].update(packed[0], n)                   # device scalars: no host read in the step
        i = 1
        if loss_predwd is not None:
            self.metrics['loss_predwd'].update(packed[i], n)
            i += 1
        if logits is not None:
            self.metrics['top1'].update(packed[i], n)
            self.metrics['top5'].update(packed[i + 1], n)
        if (self._step + 1) % max(self.log_interval, 1) == 0:
            # The update itself is guarded on the device every step; here (one host read per log interval) the count of
            # skipped updates is fetched and, as in the reference (trainer.py:240-257), a single process stops on a NaN
            # loss while data-parallel runs keep going with the poisoned steps skipped.
            self.check_finite()
        self._step += 1
        return self.metrics

    def check_finite(self):
        """Reads the device-side skip counter (one synchronisation); raises outside data parallelism if updates were
        skipped because the loss / gradients were not finite."""
        if self._fused_clip:
            self.skipped_updates = int(self._optimizer.skipped.item())
        avg = self.metrics['loss'].avg
        if (avg != avg or self.skipped_updates > 0) and not self.ddp:
            raise RuntimeError('the loss is NaN (or the gradients are not finite) at step %d: %d update(s) were skipped '
                               'on the device; unable to proceed. Restarting from the last checkpoint may help.'
                               % (self._step, self.skipped_updates))
        return self.skipped_updates
