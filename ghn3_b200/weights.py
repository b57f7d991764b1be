"""
GHN-3 checkpoint layout (the state_dict contract) and a procedural random-init generator.

The key names and shapes below are the reference's state_dict layout (reference
ghn3/nn.py:140-172 for the GHN-3 part, ghn3/graphormer.py:80-99,198-206 for the Graphormer layers; the ppuda
base-class part is pinned by ghn3/nn.py:69-88,167-169,727-733 and by the 654 365 184-parameter known answer,
examples/ghn_all_pytorch.ipynb:109).

`procedural_state_dict` produces random-init-like weights that depend only on (config, seed, tensor name) and NOT
on module construction order, so the reference (under the ppuda shim), the oracle and the CUDA path can all be
loaded with bit-identical weights without shipping a checkpoint.
"""

import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

N_PRIMITIVES = 15          # len(PRIMITIVES_DEEPNETS1M), rows of embed.weight
MAX_DEGREE = 100           # graphormer.py:196
MAX_INPUT_DIST = 1000      # graphormer.py:197
EDGE_EMBED_ROWS = 257      # graphormer.py:96


def channel_bins(num_classes):
    """Channel-size bins of ppuda's ShapeEncoder."""
    return np.unique([1, 3, num_classes] + list(range(8, 64, 8)) + list(range(64, 4096, 16)) +
                     list(range(4096, 8193, 32)))


def spatial_bins(max_shape):
    """Spatial-size bins of ppuda's ShapeEncoder."""
    return np.unique(list(range(1, max(12, max_shape[3]), 2)) + [14, 16])


def normalize_config(config):
    cfg = dict(config)
    ms = cfg['max_shape']
    if not isinstance(ms, (tuple, list)):
        s = 16 if cfg.get('num_classes', 1000) >= 1000 else 11
        ms = (ms, ms, s, s)
    cfg['max_shape'] = tuple(int(v) for v in ms)
    cfg.setdefault('num_classes', 1000)
    cfg.setdefault('heads', 8)
    cfg.setdefault('layers', 3)
    cfg.setdefault('layernorm', True)
    return cfg


def state_dict_spec(config):
    """Ordered {key: shape} of a GHN-3 state_dict (after fix_embed_layers, i.e. embeddings under gnn.0.)."""
    cfg = normalize_config(config)
    C, L, H = cfg['hid'], cfg['layers'], cfg['heads']
    ms0, ms1, s2, s3 = cfg['max_shape']
    ncls = cfg['num_classes']
    n_ch, n_s = len(channel_bins(ncls)), len(spatial_bins(cfg['max_shape']))
    max_ch = max(ms0, ms1)
    spec = OrderedDict()
    if cfg['layernorm']:
        spec['ln.weight'] = (C,)
        spec['ln.bias'] = (C,)
    spec['embed.weight'] = (N_PRIMITIVES, C)
    spec['shape_enc.embed_spatial.weight'] = (n_s + 1, C // 4)
    spec['shape_enc.embed_channel.weight'] = (n_ch + 1, C // 4)
    for l in range(L):
        p = 'gnn.%d.' % l
        spec[p + 'ln1.weight'] = (C,)
        spec[p + 'ln1.bias'] = (C,)
        spec[p + 'attn.to_qkv.weight'] = (3 * C, C)
        spec[p + 'attn.to_out.0.weight'] = (C, C)
        spec[p + 'attn.to_out.0.bias'] = (C,)
        if l == 0:
            spec[p + 'attn.edge_embed.embed.weight'] = (EDGE_EMBED_ROWS, C)
            spec[p + 'attn.proj_e.0.weight'] = (C, 2 * C)
            spec[p + 'attn.proj_e.0.bias'] = (C,)
            spec[p + 'attn.proj_e.2.weight'] = (H, C)
            spec[p + 'attn.proj_e.2.bias'] = (H,)
        spec[p + 'ln2.weight'] = (C,)
        spec[p + 'ln2.bias'] = (C,)
        spec[p + 'ff.net.0.weight'] = (4 * C, C)
        spec[p + 'ff.net.0.bias'] = (4 * C,)
        spec[p + 'ff.net.3.weight'] = (C, 4 * C)
        spec[p + 'ff.net.3.bias'] = (C,)
        if l == 0:
            spec[p + 'centrality_embed_in.weight'] = (MAX_DEGREE + 1, C)
            spec[p + 'centrality_embed_out.weight'] = (MAX_DEGREE + 1, C)
            spec[p + 'input_dist_embed.weight'] = (MAX_INPUT_DIST + 1, C)
    spec['decoder.fc.0.weight'] = (4 * C * s2 * s3, C)
    spec['decoder.fc.0.bias'] = (4 * C * s2 * s3,)
    spec['decoder.conv.0.weight'] = (8 * C, 4 * C)
    spec['decoder.conv.0.bias'] = (8 * C,)
    spec['decoder.conv.2.weight'] = (ms0 * ms1, 8 * C)
    spec['decoder.conv.2.bias'] = (ms0 * ms1,)
    spec['decoder.class_layer_predictor.1.weight'] = (ncls, ms0)
    spec['decoder.class_layer_predictor.1.bias'] = (ncls,)
    spec['decoder_1d.fc.0.weight'] = (2 * C, C)
    spec['decoder_1d.fc.0.bias'] = (2 * C,)
    spec['decoder_1d.fc.2.weight'] = (2 * max_ch, 2 * C)
    spec['decoder_1d.fc.2.bias'] = (2 * max_ch,)
    spec['bias_class.1.weight'] = (ncls, max_ch)
    spec['bias_class.1.bias'] = (ncls,)
    return spec


def num_parameters(config):
    return int(sum(int(np.prod(s)) for s in state_dict_spec(config).values()))


def sinusoid_table(rows, hid):
    """The EdgeEmbedding initial table (graphormer.py:55-65): sin/cos positional code with row 0 zeroed."""
    position = torch.arange(rows).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, hid, 2) * (-math.log(10000.0) / hid))
    pe = torch.zeros(rows, hid)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    pe[0, :] = 0
    return pe


_SMALL_INIT = ('decoder_1d.fc.2.', 'decoder.conv.2.', 'decoder.class_layer_predictor.1.')   # nn.py:167-169


def procedural_state_dict(config, seed=0):
    """Random-init-like fp32 weights as a pure function of (config, seed, key)."""
    spec = state_dict_spec(config)
    sd = OrderedDict()
    for key, shape in spec.items():
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))
        n = int(np.prod(shape))
        if key.endswith('edge_embed.embed.weight'):
            t = sinusoid_table(shape[0], shape[1])
        elif 'embed' in key:
            # nn.py:704-713: trunc_normal(std=d**-0.5) -> here a clipped normal with the same std
            d = shape[1]
            a = rng.standard_normal(n, dtype=np.float32)
            np.clip(a, -2.0, 2.0, out=a)
            t = torch.from_numpy(a).reshape(shape) * (d ** -0.5)
        elif len(shape) == 1 and ('.ln1.' in key or '.ln2.' in key or key.startswith('ln.')):
            u = rng.random(n, dtype=np.float32) * 2 - 1
            t = torch.from_numpy(u) * 0.05 + (1.0 if key.endswith('weight') else 0.0)
        else:
            fan_in = shape[1] if len(shape) == 2 else spec[key[:-len('bias')] + 'weight'][1]
            bound = 1.0 / math.sqrt(fan_in)
            small = key.startswith(_SMALL_INIT)
            if small and key.endswith('bias'):
                t = torch.zeros(shape)                     # nn.py:701
            else:
                u = rng.random(n, dtype=np.float32)
                u *= 2 * bound
                u -= bound
                if small:
                    u /= 5.0                               # nn.py:700
                t = torch.from_numpy(u).reshape(shape)
        sd[key] = t.contiguous()
    return sd


# Named GHN-3 configurations (SURVEY.md §8 header; T from train_ghn_ddp.py:17,58, XL pinned by the param count).
CONFIGS = {
    'ghn3tm8': dict(hid=64, layers=3, heads=8, max_shape=(64, 64, 16, 16), num_classes=1000, layernorm=True),
    'ghn3sm8': dict(hid=128, layers=5, heads=16, max_shape=(128, 128, 16, 16), num_classes=1000, layernorm=True),
    'ghn3lm8': dict(hid=256, layers=12, heads=16, max_shape=(256, 256, 16, 16), num_classes=1000, layernorm=True),
    'ghn3xlm16': dict(hid=384, layers=24, heads=16, max_shape=(384, 384, 16, 16), num_classes=1000, layernorm=True),
    # tiny config for fast tests (not a released model)
    'ghn3tiny': dict(hid=32, layers=2, heads=8, max_shape=(32, 32, 16, 16), num_classes=1000, layernorm=True),
}
