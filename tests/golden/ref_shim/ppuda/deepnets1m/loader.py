import torch.utils.data
MAX_NODES_BATCH = 2200
class DeepNets1M(torch.utils.data.Dataset): pass
class NetBatchSampler(torch.utils.data.BatchSampler): pass
