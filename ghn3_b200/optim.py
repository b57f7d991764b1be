"""
Optimizer of the training path: global-norm gradient clipping + AdamW in one kernel pass (ghn3_adamw) over the flat
gradient buffer the hand-written backward pass produces. Replaces `nn.utils.clip_grad_norm_` +
`torch.optim.AdamW.step` of the reference's `Trainer.update` (ghn3/trainer.py:343-379): no per-tensor norm kernels, no
host synchronisation, 7 HBM passes over the parameters' bytes instead of ~12.
"""
import numpy as np
import torch

from . import _lib as L
from .train import flat_layout

CHUNK = 8192


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, ghn, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        params, order, offs, early = flat_layout(ghn)
        self._ghn_ref, self._early = ghn, early
        super().__init__(order, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                     max_grad_norm=max_grad_norm))
        dev = order[0].device
        if dev.type != 'cuda':
            raise RuntimeError('ghn3_b200.FusedAdamW: the GHN must be on a CUDA device')
        for p in order:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('ghn3_b200.FusedAdamW: parameters must be contiguous fp32 tensors')
        self._order, self._offs, self._dev = order, offs, dev
        total = int(offs[-1])
        self._total = total
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._own_grads = None
        numels = np.array([p.numel() for p in order], dtype=np.int64)
        chunks = (numels + CHUNK - 1) // CHUNK
        chunk0 = np.concatenate([[0], np.cumsum(chunks)[:-1]]).astype(np.int64)
        self._n_chunks = int(chunks.sum())
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self._tables = dict(offsets=to_dev(offs[:-1]), numels=to_dev(numels), chunk0=to_dev(chunk0),
                            chunk_tensor=to_dev(np.repeat(np.arange(len(order), dtype=np.int32), chunks)))
        self._ptrs_host, self._ptrs = None, None
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        # non-finite guard (ghn3_adamw_args.skipped): updates skipped on the device because |g|^2 or the loss was not
        # finite; `loss_for_guard` may be set to a device scalar before step()
        self.skipped = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_for_guard = None
        self._step = 0
        self._checked = False
        self._sync = None                          # GradSync(shard=True): sharded step, see enable_sharding
        self._pflat = None
        # flat start of every chunk (host): the chunks that intersect a shard are found by bisection
        self._chunk_start = np.repeat(offs[:-1], chunks) + \
            (np.arange(self._n_chunks, dtype=np.int64) - np.repeat(chunk0, chunks)) * CHUNK

    def enable_sharding(self, sync):
        """Data parallelism with a reduce-scattered gradient (train.GradSync(shard=True)): this rank updates only its
        slice of each flat region -- AdamW over 1/N of the 2.6 GB of parameters and moments instead of all of them on
        every rank -- and the updated parameters are all-gathered. The parameters become views of ONE flat fp32 buffer
        (same layout as the gradient buffer) so that the all-gather writes them in place; the update itself is
        arithmetically the same as the replicated one (the global gradient norm is summed over the ranks)."""
        if sync is None or not getattr(sync, 'shard', False):
            return False
        self._sync = sync
        self._flatten_params()
        return True

    def _flatten_params(self):
        pflat = torch.zeros(self._total, dtype=torch.float32, device=self._dev)
        with torch.no_grad():
            for p, o in zip(self._order, self._offs[:-1]):
                v = pflat[int(o):int(o) + p.numel()].view(p.shape)
                v.copy_(p.detach())
                p.data = v
        self._pflat = pflat
        self._ptrs_host = None
        ghn = self._ghn_ref
        ghn._dev = None                            # device weight copies alias the parameters' storage: rebuild
        ghn.__dict__['_param_list'] = None

    def _params_are_flat(self):
        base = self._pflat.data_ptr()
        n = len(self._order)
        return all(self._order[i].data_ptr() == base + int(self._offs[i]) * 4 for i in (0, n // 2, n - 1))

    def _chunk_range(self, lo, hi):
        c0 = int(np.searchsorted(self._chunk_start, lo, side='right')) - 1
        c1 = int(np.searchsorted(self._chunk_start, hi, side='left'))
        return max(c0, 0), max(c1 - max(c0, 0), 0)

    def gather_state(self):
        """Sharded mode: all-gathers the moment slices so that every rank holds the full optimizer state (collective;
        called by Trainer.save on every rank before rank 0 writes the checkpoint)."""
        sync = self._sync
        if sync is None:
            return
        for lo, hi in ((0, self._early), (self._early, self._total)):
            if hi > lo:
                a, b = sync.shard_of(lo, hi)
                for buf in (self.exp_avg, self.exp_avg_sq):
                    sync.dist.all_gather_into_tensor(buf[lo:hi], buf[a:b], group=sync.group)

    def _param_table(self):
        ptrs = [p.data_ptr() for p in self._order]
        if ptrs != self._ptrs_host:
            self._ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self._dev)
            self._ptrs_host = ptrs
        return self._ptrs

    def _flat_grads(self):
        """The flat gradient buffer: the backward pass's own buffer when every p.grad is still the view it handed
        out (the normal case, zero copies); otherwise the gradients are gathered into a private flat buffer."""
        g0 = self._order[0].grad
        if g0 is not None:
            base = g0.data_ptr()
            # every gradient handed out by the backward pass is a slice of ONE buffer at the known offsets; the full
            # check of all tensors runs on the first step, afterwards three sentinels (first, middle, last) suffice
            idx = range(len(self._order)) if not self._checked else (0, len(self._order) // 2, len(self._order) - 1)
            if all(self._order[i].grad is not None and self._order[i].grad.is_contiguous() and
                   self._order[i].grad.data_ptr() == base + int(self._offs[i]) * 4 for i in idx):
                self._checked = True
                return base, None
            self._checked = False
        if self._own_grads is None:
            self._own_grads = torch.zeros(self._total, dtype=torch.float32, device=self._dev)
        flat = self._own_grads
        flat.zero_()
        views = [flat[int(o):int(o) + p.numel()].view(p.shape) for p, o in zip(self._order, self._offs[:-1])]
        have = [(v, p.grad) for v, p in zip(views, self._order) if p.grad is not None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        return flat.data_ptr(), flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self._step += 1
        b1, b2 = g['betas']
        base, keep = self._flat_grads()
        t = self._tables
        a = L.AdamWArgs(params=self._param_table().data_ptr(), grads=base, exp_avg=self.exp_avg.data_ptr(),
                        exp_avg_sq=self.exp_avg_sq.data_ptr(), offsets=t['offsets'].data_ptr(),
                        numels=t['numels'].data_ptr(), chunk0=t['chunk0'].data_ptr(),
                        chunk_tensor=t['chunk_tensor'].data_ptr(), n_chunks=self._n_chunks, total=self._total,
                        lr=g['lr'], beta1=b1, beta2=b2, eps=g['eps'], weight_decay=g['weight_decay'],
                        bias_correction1=1.0 - b1 ** self._step, bias_correction2=1.0 - b2 ** self._step,
                        max_norm=float(g['max_grad_norm'] or 0.0), sumsq=self._sumsq.data_ptr(),
                        skipped=self.skipped.data_ptr())
        lg = self.loss_for_guard
        if lg is not None:
            lg = lg.detach().reshape(-1)[:1].float().contiguous()
            a.loss = lg.data_ptr()
            self._loss_keep = lg
            self.loss_for_guard = None
        sync = self._sync
        if sync is None:
            L.call('adamw', a, L.current_stream())
            return loss
        # ---- sharded step: |g|^2 of this rank's slices -> sum over ranks -> AdamW on the slices -> all-gather ----
        if not self._params_are_flat():            # .to() / load_state_dict replaced the storages
            self._flatten_params()
            a.params = self._param_table().data_ptr()
        regions = [(lo, hi) for lo, hi in ((0, self._early), (self._early, self._total)) if hi > lo]
        shards = [sync.shard_of(lo, hi) for lo, hi in regions]
        for k, (lo, hi) in enumerate(shards):
            a.range_lo, a.range_hi, a.sumsq_ready = lo, hi, (-1 if k == 0 else -2)
            L.call('adamw', a, L.current_stream())
        sync.dist.all_reduce(self._sumsq, group=sync.group)
        for k, (lo, hi) in enumerate(shards):
            a.range_lo, a.range_hi = lo, hi
            a.chunk_begin, a.n_chunks = self._chunk_range(lo, hi)
            a.sumsq_ready = 1 if k == 0 else 2
            L.call('adamw', a, L.current_stream())
        for (lo, hi), (s0, s1) in zip(regions, shards):
            sync.dist.all_gather_into_tensor(self._pflat[lo:hi], self._pflat[s0:s1], group=sync.group)
        return loss

    def state_dict(self):
        """Optimizer state for checkpoints (reference trainer.py:419-426 stores `optimizer.state_dict()`): the two flat
        moment buffers, in the [decoder | rest] layout of ghn3_b200.train.flat_layout, and the step count."""
        if self._sync is not None:
            self.gather_state()
        d = super().state_dict()
        d['flat'] = {'exp_avg': self.exp_avg.detach().clone(), 'exp_avg_sq': self.exp_avg_sq.detach().clone(),
                     'step': self._step}
        return d

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop('flat', None)
        super().load_state_dict(state_dict)
        if flat is not None:
            if flat['exp_avg'].numel() != self._total:
                raise ValueError('FusedAdamW: checkpoint moments have %d elements, expected %d'
                                 % (flat['exp_avg'].numel(), self._total))
            self.exp_avg.copy_(flat['exp_avg'])
            self.exp_avg_sq.copy_(flat['exp_avg_sq'])
            self._step = int(flat['step'])

    def grad_norm(self):
        """Global gradient norm seen by the last step (device tensor, float64) when clipping is enabled."""
        return self._sumsq.sqrt()
