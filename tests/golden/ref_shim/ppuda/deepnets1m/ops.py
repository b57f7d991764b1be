import torch, torch.nn as nn
def parse_op_ks(op):
    import re
    m = re.match(r'^(.*)_(\d+)x(\d+)$', op)
    return (m.group(1), int(m.group(2))) if m else (op, 3)
class PosEnc(nn.Module):
    def __init__(self, C, ks):
        super().__init__(); self.weight = nn.Parameter(torch.randn(1, C, ks, ks))
    def forward(self, x): return x + self.weight
