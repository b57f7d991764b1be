"""Race hunt for the two-lane / CUDA-graph / multi-stream prediction path: many predictions over alternating batch plans
in every launch mode, each compared with reference parameters from a one-lane, graph-free, serial GHN."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from ghn3_b200 import GHN3, Graph, GraphBatch
from ghn3_b200.weights import CONFIGS, procedural_state_dict
from tests import helpers as H

dev = torch.device('cuda:0')
records = B.load_records()
cfgname = sys.argv[1] if len(sys.argv) > 1 else 'ghn3sm8'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = CONFIGS[cfgname]
plans = [['vit_b_16', 'convnext_tiny'], ['resnet50'], ['efficientnet_b0', 'squeezenet1_1', 'mobilenet_v3_small']]
worst_all = 0.0
for dtype, tol in (('tf32', 2e-4), ('bf16', 3e-2)):
    sd = procedural_state_dict(cfg, 0)

    def make(**kw):
        g = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype=dtype)
        g.load_state_dict(sd)
        g = g.to(dev).eval()
        for k, v in kw.items():
            setattr(g, k, v)
        return g
    ref_ghn = make(decoder_lanes=False, cuda_graphs=False)
    refs = []
    with torch.no_grad():
        for archs in plans:
            ms = [B.build_model(a).to(dev) for a in archs]
            for m in ms:
                for p in m.parameters():
                    p.zero_()
            ref_ghn(ms, [Graph.from_record(records[a]) for a in archs])
            torch.cuda.synchronize()
            refs.append([[p.detach().clone() for p in m.parameters()] for m in ms])
    del ref_ghn
    for mode in ('serial', 'overlap1', 'overlap2', 'overlap4'):
        for graphs_on in (True, False):
            ghn = make(cuda_graphs=graphs_on)
            if mode != 'serial':
                ghn.overlap_scatter = True
                ghn.pipeline_depth = int(mode[-1])
            models = [[B.build_model(a).to(dev) for a in archs] for archs in plans]
            gl = [[Graph.from_record(records[a]) for a in archs] for archs in plans]
            worst = 0.0
            nbad = [0]
            with torch.no_grad():
                for it in range(iters):
                    k = (it * 7 + it // 3) % len(plans)
                    for m in models[k]:
                        for p in m.parameters():
                            p.zero_()
                    ghn(models[k], GraphBatch(gl[k], dense=True).to_device(dev))
                    if it % 3 != 2:                 # sometimes let several predictions pile up before checking
                        continue
                    ghn.flush_all()
                    torch.cuda.synchronize()
                    for m, rp in zip(models[k], refs[k]):
                        for (n, p), r in zip(m.named_parameters(), rp):
                            if n.endswith('class_token') or 'pos_embedding' in n or not bool(r.any()):
                                continue            # fresh random draws / parameters the GHN does not predict
                            e = H.max_rel_err(p, r)
                            worst = max(worst, e)
                            if e > tol:
                                nbad[0] += 1
                                if nbad[0] <= 5:
                                    print('MISMATCH', dtype, mode, graphs_on, it, plans[k], n, e, flush=True)
            print('%s %-8s graphs=%-5s worst %.2e mismatches %d' % (dtype, mode, graphs_on, worst, nbad[0]), flush=True)
            worst_all = max(worst_all, worst)
            del ghn
print('done')
