"""Where does the host spend its time in one end-to-end step of bench.py (pack graphs -> H2D -> SPD -> ghn() -> norms)?
Prints host milliseconds per part (perf_counter, no device synchronisation inside the loop) and the device-side pace.
  python tools/e2e_host_profile.py [--depth 4] [--steps 200]
"""
import argparse
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--depth', type=int, default=4)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--cprofile', action='store_true', help='cProfile of the ghn() calls (top 35 by cumulative time)')
    args = ap.parse_args()
    import torch
    import bench as B
    from ghn3_b200 import GHN3, Graph, GraphBatch
    from ghn3_b200.weights import CONFIGS, procedural_state_dict
    dev = torch.device('cuda', 0)
    cfg = CONFIGS['ghn3xlm16']
    records = B.load_records()
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
    ghn.load_state_dict(procedural_state_dict(cfg, 0))
    ghn = ghn.to(dev).eval()
    models = [B.build_model(a).to(dev) for a in B.WORKLOAD_ARCHS]
    graphs = [Graph.from_record(records[a]) for a in B.WORKLOAD_ARCHS]
    ghn.overlap_scatter = True
    ghn.pipeline_depth = args.depth
    nb = args.depth + 1
    bufs = [torch.empty(len(models), dtype=torch.float64).pin_memory() for _ in range(nb)]
    evs = [torch.cuda.Event() for _ in range(nb)]
    t = dict(pack=0.0, to_device=0.0, ghn=0.0, norms=0.0, wait=0.0)

    def run(n, acc):
        pending = []
        for k in range(n):
            t0 = time.perf_counter()
            gb = GraphBatch(graphs, dense=True)
            t1 = time.perf_counter()
            b = gb.to_device(dev)
            t2 = time.perf_counter()
            ghn(models, b)
            t3 = time.perf_counter()
            with torch.cuda.stream(ghn.result_stream()):
                bufs[k % nb].copy_(ghn.param_norms(models), non_blocking=True)
                evs[k % nb].record()
            t4 = time.perf_counter()
            pending.append(k % nb)
            if len(pending) > args.depth:
                evs[pending.pop(0)].synchronize()
            t5 = time.perf_counter()
            if acc:
                t['pack'] += t1 - t0
                t['to_device'] += t2 - t1
                t['ghn'] += t3 - t2
                t['norms'] += t4 - t3
                t['wait'] += t5 - t4
        for j in pending:
            evs[j].synchronize()

    with torch.no_grad():
        run(10, False)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        run(args.steps, True)
        ghn.flush_all()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
    if args.cprofile:
        import cProfile
        import pstats
        b = GraphBatch(graphs, dense=True).to_device(dev)
        pr = cProfile.Profile()
        with torch.no_grad():
            pr.enable()
            for _ in range(args.steps):
                ghn(models, b)
            pr.disable()
        torch.cuda.synchronize()
        st = pstats.Stats(pr)
        st.sort_stats('cumulative').print_stats(35)
    n = args.steps
    print('e2e wall %.3f ms/step (depth %d); host ms/step: %s' % (
        (w1 - w0) / n * 1e3, args.depth, ', '.join('%s %.3f' % (k, v / n * 1e3) for k, v in t.items())))


if __name__ == '__main__':
    main()
