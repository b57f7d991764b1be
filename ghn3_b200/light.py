"""
Parameter-free ("light") target modules for GHN training -- the role of the reference's ModuleLight classes
(ghn3/ops.py:60-101, ghn3/light_ops.py:26-337): a network whose layers hold only the SHAPES of their parameters
(`weight = [out, in, kh, kw]`) until a GHN predicts them with `keep_grads=True`, after which the attributes are the
predicted tensors (with a grad_fn). Nothing is allocated or initialised per architecture, which is what makes
sampling a fresh meta-batch every step cheap.

This is an independent, minimal implementation: `Module` below is a plain Python object tree (not an nn.Module) with
the few members the prediction / training path touches -- named_modules(), parameters(), named_parameters(),
__call__, train() / eval(), apply(), to(). Layers are functional wrappers over torch.nn.functional.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F


class Module:
    """Base of the light modules: children and (shape-or-tensor) parameters are tracked by attribute assignment."""

    def __init__(self):
        object.__setattr__(self, '_modules', OrderedDict())
        object.__setattr__(self, '_parameters', OrderedDict())
        object.__setattr__(self, 'training', True)

    def __setattr__(self, name, value):
        if isinstance(value, (Module, torch.nn.Module)):
            self._modules[name] = value
        elif name in ('weight', 'bias') and (value is None or isinstance(value, (list, tuple, torch.Tensor))):
            self._parameters[name] = value
        object.__setattr__(self, name, value)

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, *args, **kwargs):                       # pragma: no cover
        raise NotImplementedError

    def add_module(self, name, module):
        setattr(self, name, module)

    def children(self):
        return iter(self._modules.values())

    def named_modules(self, memo=None, prefix=''):
        memo = set() if memo is None else memo
        if id(self) in memo:
            return
        memo.add(id(self))
        yield prefix, self
        for name, m in self._modules.items():
            sub = prefix + ('.' if prefix else '') + name
            if isinstance(m, Module):
                yield from m.named_modules(memo, sub)
            else:                                             # a plain nn.Module leaf (e.g. nn.ReLU)
                for n2, m2 in m.named_modules(prefix=sub):
                    yield n2, m2

    def modules(self):
        for _, m in self.named_modules():
            yield m

    def named_parameters(self, prefix='', recurse=True):
        """Yields only what a GHN has predicted so far (tensors); shape placeholders are skipped."""
        for name, m in self.named_modules():
            if not isinstance(m, Module):
                for n2, p in m.named_parameters(recurse=False):
                    yield (name + '.' if name else '') + n2, p
                continue
            for key, p in m._parameters.items():
                if isinstance(p, torch.Tensor):
                    yield (name + '.' if name else '') + key, p

    def parameters(self, recurse=True):
        for _, p in self.named_parameters():
            yield p

    def apply(self, fn):
        for m in list(self.modules()):
            fn(m)
        return self

    def train(self, mode=True):
        for m in self.modules():
            if isinstance(m, Module):
                object.__setattr__(m, 'training', mode)
            else:
                m.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, *args, **kwargs):
        """Predicted tensors already live on the GHN's device; shapes have no device."""
        return self

    def cuda(self, *args, **kwargs):
        return self


class Sequential(Module):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            setattr(self, str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            x = m(x)
        return x

    def __iter__(self):
        return iter(self._modules.values())

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]


class ModuleList(Module):
    def __init__(self, mods=()):
        super().__init__()
        for m in mods:
            self.append(m)

    def append(self, m):
        setattr(self, str(len(self._modules)), m)
        return self

    def __iter__(self):
        return iter(self._modules.values())

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class Conv2d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.groups = in_channels, out_channels, groups
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.weight = [out_channels, in_channels // groups, *self.kernel_size]
        self.bias = [out_channels] if bias else None

    def forward(self, x):
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class Linear(Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = [out_features, in_features]
        self.bias = [out_features] if bias else None

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class BatchNorm2d(Module):
    """Batch statistics only (the reference's light BatchNorm asserts track_running_stats=False, light_ops.py:283)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.track_running_stats = False
        self.weight = [num_features]
        self.bias = [num_features]

    def forward(self, x):
        return F.batch_norm(x, None, None, self.weight, self.bias, True, self.momentum, self.eps)


class LayerNorm(Module):
    def __init__(self, normalized_shape, eps=1e-5):
        super().__init__()
        self.normalized_shape = (normalized_shape,) if isinstance(normalized_shape, int) else tuple(normalized_shape)
        self.eps = eps
        self.weight = list(self.normalized_shape)
        self.bias = list(self.normalized_shape)

    def forward(self, x):
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)


class _Fn(Module):
    """Parameter-free layer."""

    def __init__(self, fn):
        super().__init__()
        object.__setattr__(self, '_fn', fn)

    def forward(self, x):
        return self._fn(x)


def ReLU(inplace=False):
    return _Fn(F.relu)


def GELU():
    return _Fn(F.gelu)


def Identity():
    return _Fn(lambda x: x)


def MaxPool2d(kernel_size, stride=None, padding=0):
    return _Fn(lambda x: F.max_pool2d(x, kernel_size, stride, padding))


def AvgPool2d(kernel_size, stride=None, padding=0, count_include_pad=True):
    return _Fn(lambda x: F.avg_pool2d(x, kernel_size, stride, padding, count_include_pad=count_include_pad))


class _TorchNamespace:
    """The same layer vocabulary backed by ordinary torch.nn modules."""
    Module, Sequential, ModuleList = torch.nn.Module, torch.nn.Sequential, torch.nn.ModuleList
    Conv2d, Linear, BatchNorm2d, LayerNorm = torch.nn.Conv2d, torch.nn.Linear, torch.nn.BatchNorm2d, torch.nn.LayerNorm
    ReLU, GELU, Identity, MaxPool2d, AvgPool2d = (torch.nn.ReLU, torch.nn.GELU, torch.nn.Identity, torch.nn.MaxPool2d,
                                                  torch.nn.AvgPool2d)
    light = False


class _LightNamespace:
    Module, Sequential, ModuleList = Module, Sequential, ModuleList
    Conv2d, Linear, BatchNorm2d, LayerNorm = Conv2d, Linear, BatchNorm2d, LayerNorm
    ReLU, GELU, Identity = staticmethod(ReLU), staticmethod(GELU), staticmethod(Identity)
    MaxPool2d, AvgPool2d = staticmethod(MaxPool2d), staticmethod(AvgPool2d)
    light = True


TORCH, LIGHT = _TorchNamespace, _LightNamespace
