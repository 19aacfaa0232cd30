import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpvsg_b200 import ops
from oracle import m2f as om
print('cpu capability', torch.backends.cpu.get_cpu_capability(), 'threads', torch.get_num_threads())
os.system("lscpu | grep -i 'model name' | head -1")
pe = ops.sine_pe(23, 40, 'cuda').cpu()
ref = om.sine_pe_2d(1, 23, 40)[0].flatten(1).t()
err = (pe - ref).abs()
i = int(err.argmax()); r, c = divmod(i, 256)
print('max err', float(err.max()), 'at token', r, 'channel', c, 'gpu', float(pe[r, c]), 'cpu', float(ref[r, c]))
# float64 truth
h, w = 23, 40
y, x = divmod(r, w)
nf = 128
k = c if c < nf else c - nf
e = ((y + 1) if c < nf else (x + 1)) / ((h if c < nf else w) + 1e-6) * 2 * math.pi
dt = 10000 ** (2 * (k // 2) / nf)
truth = math.cos(e / dt) if k & 1 else math.sin(e / dt)
print('float64 truth', truth)
dim_t = torch.arange(nf, dtype=torch.float32)
dim_t32 = 10000 ** (2 * (dim_t // 2) / nf)
print('dim_t[k] f32', float(dim_t32[k]), 'f64', dt)
