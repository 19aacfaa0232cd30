import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpvsg_b200 import lib, ops
dev = torch.device('cuda'); g = torch.Generator().manual_seed(0)
B, H, Lq, Lk, D = (int(x) for x in sys.argv[1:6])
E = H * D
q = torch.randn(B, Lq, E, generator=g).to(dev)
k = ops.Split(*ops.split_bf16(torch.randn(B, Lk, E, generator=g).to(dev)))
v = ops.Split(*ops.split_bf16(torch.randn(B, Lk, E, generator=g).to(dev)))
for _ in range(3):
    ops.attention(q, k, v, H)
torch.cuda.synchronize()
