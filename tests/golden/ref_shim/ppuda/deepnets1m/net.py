import torch, torch.nn as nn
def _is_none(mod):
    for n, p in mod.named_modules():
        if hasattr(p, 'weight') and p.weight is None: return True
    return False
def drop_path(x, p): return x
class AuxiliaryHeadImageNet(nn.Module): pass
class AuxiliaryHeadCIFAR(nn.Module): pass
class Network(nn.Module): pass
def get_cell_ind(param_name, layers=1):
    if param_name.find('cells.') >= 0:
        pos1 = len('cells.'); pos2 = pos1 + param_name[pos1:].find('.')
        return int(param_name[pos1:pos2])
    if param_name.startswith(('classifier','auxiliary')): return layers - 1
    if layers == 1 or param_name.startswith(('stem','pos_enc')): return 0
    return None
def named_layered_modules(model):
    if hasattr(model, 'module'): model = model.module
    layers = model._n_cells if hasattr(model, '_n_cells') else 1
    lm = [{} for _ in range(layers)]
    for name, m in model.named_modules():
        entries = []
        for attr, suffix, is_w in [('weight','.weight',True),('bias','.bias',False),('in_proj_weight','.in_proj_weight',True),
                                   ('in_proj_bias','.in_proj_bias',False),('pos_embedding','.pos_embedding.weight',True)]:
            p = getattr(m, attr, None)
            if p is None or isinstance(p, bool): continue
            if not isinstance(p, (torch.Tensor, list, tuple)): continue
            entries.append((name + suffix, p, is_w))
        if not entries: continue
        ci = get_cell_ind(name, layers); ci = 0 if ci is None else ci
        for key, p, is_w in entries:
            lm[ci][key] = {'param_name': key, 'module': m, 'is_w': is_w, 'sz': tuple(p) if isinstance(p,(list,tuple)) else p.shape}
    return lm
