"""Per-phase timing of ghn3_graphormer_fused from its in-kernel trace (thread 64 of every CTA)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ghn3_b200 import _lib as L
from tests.test_fused_gpu import _setup

C, Hh, layers = 384, 16, 24
archs = sys.argv[1].split(',') if len(sys.argv) > 1 else ['vit_b_16', 'convnext_base']
recs, pack, per, stack, x0, lut, fg = _setup(C, Hh, layers, archs, seed=3)
lib = L.load()
for _ in range(5):
    fg.run()
torch.cuda.synchronize()
buf = torch.zeros(148, 1024, 3, dtype=torch.int64, device='cuda')
lib.ghn3_debug_fused_trace(ctypes.c_void_p(buf.data_ptr()))
fg.run()
torch.cuda.synchronize()
lib.ghn3_debug_fused_trace(ctypes.c_void_p(0))
t = buf.cpu().numpy()
names = ['qkv', 'attn', 'proj', 'ff1', 'ff2']
gemm_ph = ['wait deps', 'build act', 'wait mma', 'epilogue', 'signal']
attn_ph = ['wait deps', 'lut', 'stage kv', 'compute', 'merge+store', 'signal']
gt0 = min(int(t[c, 0, 2]) for c in range(148) if t[c, 0, 0] == -1)
end = max(int(t[c, :, 2].max()) for c in range(148))
fr = []
for c in range(148):
    n = int((t[c, :512, 2] != 0).sum())
    if n > 2:
        fr.append((int(t[c, n - 1, 1]) - int(t[c, 0, 1])) / max(int(t[c, n - 1, 2]) - int(t[c, 0, 2]), 1))
print('SM clock from clock64 / globaltimer: %.3f GHz (min %.3f, max %.3f)' % (np.mean(fr), np.min(fr), np.max(fr)))
print('kernel span by globaltimer: %.1f us' % ((end - gt0) / 1e3))
# per stage (layers >= 2 to skip warm-up effects): mean phase durations in ns (globaltimer) over CTAs
acc = {}
stage_span = {}
for c in range(148):
    rec = t[c]
    n = int((rec[:, 2] != 0).sum())
    i = 0
    while i < n:
        tag = int(rec[i, 0])
        if tag < 0:
            i += 1
            continue
        gs, ph = tag >> 8, tag & 255
        if ph != 0:
            i += 1
            continue
        j = i + 1
        while j < n and (int(rec[j, 0]) >> 8) == gs and (int(rec[j, 0]) & 255) != 0 and int(rec[j, 0]) >= 0:
            j += 1
        s = gs % 5
        if gs >= 10:
            for a, b in zip(range(i, j - 1), range(i + 1, j)):
                key = (s, int(rec[b, 0]) & 255)
                acc.setdefault(key, []).append(int(rec[b, 2]) - int(rec[a, 2]))
        lo, hi = stage_span.get(gs, (1 << 62, 0))
        stage_span[gs] = (min(lo, int(rec[i, 2])), max(hi, int(rec[j - 1, 2])))
        i = j
for s in range(5):
    ph = attn_ph if s == 1 else gemm_ph
    parts = []
    for k in range(1, len(ph) + 1):
        v = acc.get((s, k))
        if v:
            parts.append('%s %.2f (max %.2f)' % (ph[k - 1], np.mean(v) / 1e3, np.max(v) / 1e3))
    print('%-5s per tile, us: %s' % (names[s], ' | '.join(parts)))
# MMA lane records (second half of every CTA's buffer): 16 = activations seen, 17 = first weight block seen, 18 = issued
mm = {}
for c in range(148):
    comp = {}
    for i in range(512):
        tag = int(t[c, i, 0])
        if t[c, i, 2] != 0 and tag >= 0:
            comp[tag] = int(t[c, i, 2])
    for i in range(512, 1024):
        tag = int(t[c, i, 0])
        if t[c, i, 2] == 0 or tag < 0:
            continue
        gs, ph = tag >> 8, tag & 255
        if gs < 10:
            continue
        mm.setdefault((gs % 5, ph), []).append((c, gs, int(t[c, i, 2])))
    t[c, 0, 0] = t[c, 0, 0]
for s_ in (0, 2, 3, 4):
    rows = {}
    for ph in (16, 17, 18):
        for c, gs, tt in mm.get((s_, ph), []):
            rows.setdefault((c, gs), {})[ph] = tt
    d = {'act->mma': [], 'weights': [], 'issue': [], 'exec+wake': []}
    for (c, gs), r in rows.items():
        comp = {int(t[c, i, 0]): int(t[c, i, 2]) for i in range(512) if t[c, i, 2] != 0 and (int(t[c, i, 0]) >> 8) == gs}
        c2, c3 = comp.get((gs << 8) | 2), comp.get((gs << 8) | 3)
        if 16 in r and 17 in r and 18 in r and c2 and c3:
            d['act->mma'].append(r[16] - c2); d['weights'].append(r[17] - r[16]); d['issue'].append(r[18] - r[17])
            d['exec+wake'].append(c3 - r[18])
    print('%-5s mma lane, us (first tile of a CTA only is exact): %s' % (names[s_], ' | '.join(
        '%s %.2f' % (k, np.mean(v) / 1e3) for k, v in d.items() if v)))
print('stage windows (first start .. last end over all CTAs), layer 2..4:')
for gs in range(10, 25):
    lo, hi = stage_span[gs]
    print('  layer %d %-5s start +%.2f us  end +%.2f us  (%.2f us)' % (gs // 5, names[gs % 5], (lo - gt0) / 1e3,
                                                                       (hi - gt0) / 1e3, (hi - lo) / 1e3))
