import numpy as np
def capacity(model, is_grad=True):
    c, n = 0, 0
    for p in model.parameters():
        if p is not None and hasattr(p,'numel'): c += 1; n += p.numel()
    return c, int(n)
def rand_choice(x, up_to=None): return np.random.choice(x[:up_to])
class AvgrageMeter: pass
def accuracy(*a, **k): pass
def init(m, **k): return m
def infer(*a, **k): pass
def adjust_net(m, **k): return m
