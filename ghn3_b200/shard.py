"""
Multi-GPU sharding of the inference path: architectures are independent units (no cross-graph term anywhere in
reference ghn3/nn.py:248-328), so they are partitioned across ranks with NO data-path collective. The assignment is
longest-processing-time-first greedy on a per-architecture cost, the idea of the reference's (dead-code) balancer
GraphBatch._sort_by_nodes (ghn3/graph.py:187-241), here on a cost that reflects the CUDA path:
    cost = bytes written into the target parameters + decoder weight bytes streamed + c * N^2.
"""
import heapq


def architecture_cost(n_nodes, n_params, decoder_rows=0):
    return 4.0 * n_params + 6144.0 * decoder_rows + 64.0 * n_nodes * n_nodes


def shard_lpt(costs, n_ranks):
    """Returns n_ranks lists of indices into `costs`; deterministic (ties broken by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    out = [[] for _ in range(n_ranks)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return [sorted(s) for s in out]


def my_shard(costs, rank=None, world=None):
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return shard_lpt(costs, world)[rank]
