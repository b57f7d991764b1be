"""
DeepNets-1M-style architecture sampler for the training path (BASELINE config 5).

The reference trains on architectures stored in the DeepNets-1M HDF5 file (ghn3/deepnets1m.py:84-147), which were
produced by ppuda's NetGenerator and are instantiated as `Network` / `NetworkLight` (ghn3/ops.py:306-569). Neither the
file nor ppuda is available offline, so this module samples architectures from the same design space directly:
DARTS-format genotypes over the primitives of ghn3/ops.py:291-304, 4-18 cells, C in 32..128 step 16, two stem types,
1-2 fully connected layers, optional global pooling, BatchNorm. The networks are ordinary `nn.Module`s; their graphs
come from the host tracer (ghn3_b200/tracer.py) like those of any other model.

This is a from-scratch generator, not a port of the reference's classes: module names and the exact wiring of the
stems differ, so graphs are DeepNets-1M-*style*, not bit-identical to the released dataset.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

OPS = ('none', 'skip_connect', 'max_pool', 'avg_pool', 'conv', 'sep_conv', 'dil_conv', 'cse')
KERNELS = {'max_pool': (3,), 'avg_pool': (3,), 'conv': (1, 3, 5), 'sep_conv': (3, 5), 'dil_conv': (3, 5)}
CHANNELS = tuple(range(32, 129, 16))            # deepnets1m.py:122-128
FC_DIMS = tuple(range(64, 257, 64))


def make_classes(ns):
    """The cell-network classes over a layer vocabulary `ns` (ghn3_b200.light.TORCH: ordinary nn.Modules with real
    parameters; ghn3_b200.light.LIGHT: parameter-free modules holding shapes, reference ghn3/ops.py:60-101). Both
    variants have identical module names, hence identical graphs and node -> parameter mappings."""

    def relu_conv_bn(c_in, c_out, k=1, stride=1):
        return ns.Sequential(ns.ReLU(), ns.Conv2d(c_in, c_out, k, stride, k // 2, bias=False), ns.BatchNorm2d(c_out))

    def sep_conv(c_in, c_out, k, stride):
        """depthwise-separable convolution applied twice (DARTS)"""
        return ns.Sequential(
            ns.ReLU(), ns.Conv2d(c_in, c_in, k, stride, k // 2, groups=c_in, bias=False),
            ns.Conv2d(c_in, c_in, 1, bias=False), ns.BatchNorm2d(c_in),
            ns.ReLU(), ns.Conv2d(c_in, c_in, k, 1, k // 2, groups=c_in, bias=False),
            ns.Conv2d(c_in, c_out, 1, bias=False), ns.BatchNorm2d(c_out))

    def dil_conv(c_in, c_out, k, stride):
        return ns.Sequential(
            ns.ReLU(), ns.Conv2d(c_in, c_in, k, stride, k - k % 2, dilation=2, groups=c_in, bias=False),
            ns.Conv2d(c_in, c_out, 1, bias=False), ns.BatchNorm2d(c_out))

    class FactorizedReduce(ns.Module):
        """stride-2 skip connection: two offset 1x1 stride-2 convolutions, concatenated"""

        def __init__(self, c_in, c_out):
            super().__init__()
            self.conv_1 = ns.Conv2d(c_in, c_out // 2, 1, 2, bias=False)
            self.conv_2 = ns.Conv2d(c_in, c_out - c_out // 2, 1, 2, bias=False)
            self.bn = ns.BatchNorm2d(c_out)

        def forward(self, x):
            x = F.relu(x)
            y = F.pad(x, (0, 1, 0, 1))[:, :, 1:, 1:]
            return self.bn(torch.cat([self.conv_1(x), self.conv_2(y)], 1))

    class Zero(ns.Module):
        def __init__(self, stride):
            super().__init__()
            self.stride = stride

        def forward(self, x):
            return (x if self.stride == 1 else x[:, :, ::self.stride, ::self.stride]) * 0.0

    class ChannelSE(ns.Module):
        """squeeze-and-excitation over channels ('cse')"""

        def __init__(self, c, stride):
            super().__init__()
            self.fc1 = ns.Linear(c, max(c // 2, 4))
            self.fc2 = ns.Linear(max(c // 2, 4), c)
            self.stride = stride

        def forward(self, x):
            if self.stride > 1:
                x = F.avg_pool2d(x, self.stride)
            s = torch.sigmoid(self.fc2(F.relu(self.fc1(x.mean((2, 3))))))
            return x * s[:, :, None, None]

    class SelfAttention(ns.Module):
        def __init__(self, dim, heads):
            super().__init__()
            self.heads = heads
            self.to_qkv = ns.Linear(dim, dim * 3, bias=False)                   # reference graphormer.py:89
            self.to_out = ns.Sequential(ns.Linear(dim, dim), ns.Identity())     # reference graphormer.py:92

        def forward(self, x):
            B, N, C = x.shape
            q, k, v = self.to_qkv(x).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
            return self.to_out(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C))

    class FeedForward(ns.Module):
        def __init__(self, dim, hidden):
            super().__init__()
            self.net = ns.Sequential(ns.Linear(dim, hidden), ns.GELU(), ns.Identity(), ns.Linear(hidden, dim),
                                     ns.Identity())                              # reference graphormer.py:38-44

        def forward(self, x):
            return self.net(x)

    class MSA(ns.Module):
        """'msa' primitive (reference ghn3/ops.py:302: TransformerLayer(C, stride=s), graphormer.py:144-248 without
        edges): pre-LN self-attention + MLP (ratio 1) over the H*W positions of a feature map."""

        def __init__(self, c, stride, heads=8):
            super().__init__()
            self.stride = stride
            self.ln1 = ns.LayerNorm(c)
            self.attn = SelfAttention(c, heads)
            self.ln2 = ns.LayerNorm(c)
            self.ff = FeedForward(c, c)

        def forward(self, x):
            B, C, H, W = x.shape
            x = x.reshape(B, C, H * W).permute(0, 2, 1)
            x = x + self.attn(self.ln1(x))
            x = x + self.ff(self.ln2(x))
            x = x.permute(0, 2, 1).reshape(B, C, H, W)
            return x if self.stride == 1 else x[:, :, ::self.stride, ::self.stride]

    def make_op(name, k, c, stride):
        if name == 'none':
            return Zero(stride)
        if name == 'skip_connect':
            return ns.Identity() if stride == 1 else FactorizedReduce(c, c)
        if name == 'max_pool':
            return ns.MaxPool2d(k, stride, k // 2)
        if name == 'avg_pool':
            return ns.AvgPool2d(k, stride, k // 2, count_include_pad=False)
        if name == 'conv':
            return relu_conv_bn(c, c, k, stride)
        if name == 'sep_conv':
            return sep_conv(c, c, k, stride)
        if name == 'dil_conv':
            return dil_conv(c, c, k, stride)
        if name == 'cse':
            return ChannelSE(c, stride)
        if name == 'msa':
            return MSA(c, stride)
        raise ValueError(name)

    class Cell(ns.Module):
        """DARTS cell: two input states, `steps` intermediate nodes with two incoming edges each, concat of the nodes
        listed in `concat`."""

        def __init__(self, edges, concat, c_pp, c_p, c, reduction, reduction_prev):
            super().__init__()
            self.preprocess0 = FactorizedReduce(c_pp, c) if reduction_prev else relu_conv_bn(c_pp, c)
            self.preprocess1 = relu_conv_bn(c_p, c)
            self.edges, self.concat = edges, list(concat)
            self.ops = ns.ModuleList()
            for (name, k, src) in edges:
                stride = 2 if reduction and src < 2 else 1
                self.ops.append(make_op(name, k, c, stride))
            self.multiplier = len(self.concat)

        def forward(self, s0, s1):
            states = [self.preprocess0(s0), self.preprocess1(s1)]
            for i in range(0, len(self.ops), 2):
                a = self.ops[i](states[self.edges[i][2]])
                b = self.ops[i + 1](states[self.edges[i + 1][2]])
                states.append(a + b)
            return torch.cat([states[i] for i in self.concat], 1)

    class CellNet(ns.Module):
        """Stem -> n_cells cells (reduction at 1/3 and 2/3 of the depth) -> (global pool) -> fc_layers classifier."""

        def __init__(self, genotype, C=32, n_cells=8, stem_type=0, fc_layers=1, fc_dim=128, glob_avg=True,
                     num_classes=1000, **unused):
            super().__init__()
            self.net_args = dict(genotype=genotype, C=C, n_cells=n_cells, stem_type=stem_type, fc_layers=fc_layers,
                                 fc_dim=fc_dim, glob_avg=glob_avg, num_classes=num_classes)
            self._n_cells = n_cells            # read by the tracer / layered-module walk (reference graph.py:327)
            if stem_type == 0:                 # one strided stem (stride 4 overall)
                self.stem = ns.Sequential(ns.Conv2d(3, C, 3, 2, 1, bias=False), ns.BatchNorm2d(C), ns.ReLU(),
                                          ns.MaxPool2d(3, 2, 1))
                c_pp = c_p = C
            else:                              # ImageNet-style double stem
                self.stem = ns.Sequential(ns.Conv2d(3, C // 2, 3, 2, 1, bias=False), ns.BatchNorm2d(C // 2),
                                          ns.ReLU(), ns.Conv2d(C // 2, C, 3, 2, 1, bias=False), ns.BatchNorm2d(C))
                c_pp = c_p = C
            self.cells = ns.ModuleList()
            c, red_prev = C, False
            for i in range(n_cells):
                reduction = n_cells >= 3 and i in (n_cells // 3, 2 * n_cells // 3)
                if reduction:
                    c *= 2
                g = genotype['reduce' if reduction else 'normal']
                cell = Cell(g, genotype['reduce_concat' if reduction else 'normal_concat'], c_pp, c_p, c, reduction,
                            red_prev)
                self.cells.append(cell)
                c_pp, c_p, red_prev = c_p, cell.multiplier * c, reduction
            self.glob_avg = glob_avg
            feat = c_p if glob_avg else c_p * 4
            layers = []
            for _ in range(fc_layers - 1):
                layers += [ns.Linear(feat, fc_dim), ns.ReLU()]
                feat = fc_dim
            layers.append(ns.Linear(feat, num_classes))
            self.classifier = ns.Sequential(*layers)

        def forward(self, x):
            s0 = s1 = self.stem(x)
            for cell in self.cells:
                s0, s1 = s1, cell(s0, s1)
            x = F.adaptive_avg_pool2d(s1, 1 if self.glob_avg else 2)
            return self.classifier(torch.flatten(x, 1))

    return {'FactorizedReduce': FactorizedReduce, 'Zero': Zero, 'ChannelSE': ChannelSE, 'MSA': MSA, 'Cell': Cell,
            'CellNet': CellNet, 'make_op': make_op}


from .light import LIGHT, TORCH  # noqa: E402

_T, _L = make_classes(TORCH), make_classes(LIGHT)
CellNet, Cell, make_op = _T['CellNet'], _T['Cell'], _T['make_op']
CellNetLight = _L['CellNet']           # the role of the reference's NetworkLight (ghn3/ops.py:93-101,565-569)


def n_params_of(net):
    """Number of parameters of a CellNet / CellNetLight (shape placeholders count by their product)."""
    total = 0
    for m in net.modules():
        for p in getattr(m, '_parameters', {}).values():
            if p is None:
                continue
            total += int(np.prod(p)) if isinstance(p, (list, tuple)) else p.numel()
    return total


def sample_genotype(rng, steps=None, ops=OPS):
    """DARTS-format genotype: per intermediate node two (op, kernel, source state) edges; `none` at most once."""
    OPS = ops

    def cell():
        n = int(rng.integers(2, 5)) if steps is None else steps
        edges = []
        for node in range(n):
            srcs = rng.choice(node + 2, size=2, replace=False)
            for src in sorted(int(s) for s in srcs):
                name = OPS[int(rng.integers(1, len(OPS)))]
                k = int(rng.choice(KERNELS[name])) if name in KERNELS else 1
                edges.append((name, k, src))
        used = {e[2] for e in edges}
        concat = [i for i in range(2, n + 2) if i not in used] or [n + 1]
        return edges, concat
    normal, normal_concat = cell()
    reduce, reduce_concat = cell()
    return {'normal': normal, 'normal_concat': normal_concat, 'reduce': reduce, 'reduce_concat': reduce_concat}


def sample_net_args(rng, ops=OPS):
    """Hyper-parameters in the ranges the reference's training loader draws from (deepnets1m.py:113-133)."""
    n_cells = int(rng.integers(4, 19))
    if n_cells > 12:
        C = CHANNELS[0]
    elif n_cells > 10:
        C = int(rng.choice(CHANNELS[:2]))
    elif n_cells > 8:
        C = int(rng.choice(CHANNELS[:3]))
    else:
        C = int(rng.choice(CHANNELS))
    return dict(genotype=sample_genotype(rng, ops=ops), C=C, n_cells=n_cells, stem_type=int(rng.integers(0, 2)),
                fc_layers=int(rng.integers(1, 3)), fc_dim=int(rng.choice(FC_DIMS)), glob_avg=bool(rng.random() < 0.8))


class NetGenerator:
    """Deterministic stream of (network, Graph) pairs: `NetGenerator(seed).sample(n)`; the same seed gives the same
    global meta-batch on every rank.
      with_msa : adds the 'msa' primitive (ghn3/ops.py:302) to the op pool (a different random stream)
      light    : the returned networks are CellNetLight instances (no parameters until a GHN predicts them with
                 keep_grads=True); their graphs are traced from an ordinary twin with the same module names"""

    def __init__(self, seed=0, num_classes=1000, max_params=20e6, with_msa=False, light=False):
        self.rng = np.random.default_rng(seed)
        self.num_classes, self.max_params = num_classes, max_params
        self.ops = OPS + ('msa',) if with_msa else OPS
        self.light = light

    def sample_net(self):
        while True:
            args = sample_net_args(self.rng, self.ops)
            if n_params_of(CellNetLight(num_classes=self.num_classes, **args)) <= self.max_params:
                return CellNet(num_classes=self.num_classes, **args)

    def sample(self, n, device=None, input_size=64):
        """n (net, graph) pairs; graphs are traced on the host with a small input (the graph does not depend on it)."""
        from .graph import Graph
        out = []
        for _ in range(n):
            net = self.sample_net()
            net.expected_input_sz = input_size
            graph = Graph(net)
            if self.light:
                net = CellNetLight(**net.net_args)
                net.expected_input_sz = input_size
            elif device is not None:
                net = net.to(device)
            graph.net = net                    # as the reference's loader does (deepnets1m.py:139-142): GraphBatch.nets
            out.append((net, graph))
        return out
