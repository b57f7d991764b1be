"""Throughput of the Graphormer kernel chain ALONE (node features + 24 layers, no decoders, no scatter) when 1..8
independent chains run side by side on their own streams -- separates "the chains slow each other down" (launch rate,
SM slots) from "the decoders / scatter slow the chains down".
  python tools/chain_throughput.py [--steps 200]
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=200)
    args = ap.parse_args()
    import torch
    import bench as B
    from ghn3_b200 import GHN3, Graph, GraphBatch
    from ghn3_b200 import _lib as L
    from ghn3_b200.weights import CONFIGS, procedural_state_dict
    dev = torch.device('cuda', 0)
    cfg = CONFIGS['ghn3xlm16']
    records = B.load_records()
    ghn = GHN3(**cfg, weight_norm=True, ve=True, compute_dtype='bf16')
    ghn.load_state_dict(procedural_state_dict(cfg, 0))
    ghn = ghn.to(dev).eval()
    models = [B.build_model(a).to(dev) for a in B.WORKLOAD_ARCHS]
    graphs = [Graph.from_record(records[a]) for a in B.WORKLOAD_ARCHS]
    batch = GraphBatch(graphs, dense=True).to_device(dev)
    ghn.overlap_scatter = True
    ghn.pipeline_depth = 8
    with torch.no_grad():
        for _ in range(24):
            ghn(models, batch)
        ghn.flush_all()
        torch.cuda.synchronize()
    bp = list(ghn._plan_cache.values())[-1]
    progs = bp.programs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    for graphs_on in (True, False):
        for pdl in (True, False):
            L.set_programmatic_launch(pdl)
            for depth in (1, 2, 4, 8):
                for p in progs:
                    p.use_graphs = graphs_on
                def run(n):
                    for k in range(n):
                        p = progs[k % depth]
                        p.launch(0, p.i_ln, p.hi.cuda_stream, 'chain', high_priority=True)
                run(2 * depth)
                torch.cuda.synchronize()
                e0.record()
                for p in progs[:depth]:
                    p.hi.wait_event(e0)
                run(args.steps)
                for p in progs[:depth]:
                    cur.wait_stream(p.hi)
                e1.record()
                torch.cuda.synchronize()
                print('graphs %-5s pdl %-5s chains %d: %.3f ms per chain-run (%.3f ms latency each if perfectly parallel)'
                      % (graphs_on, pdl, depth, e0.elapsed_time(e1) / args.steps,
                         e0.elapsed_time(e1) / args.steps * depth), flush=True)


if __name__ == '__main__':
    main()
