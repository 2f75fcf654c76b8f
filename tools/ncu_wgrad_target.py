"""A few launches of the 3xTF32 weight-gradient kernel (MN-major operands) and of the decode-step fusion for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import decode_ops as K, gemm

gen = torch.Generator().manual_seed(0)
for rows, n, k in ((20 * 5440, 256, 256), (20 * 5440, 1024, 256)):
    g = torch.randn(rows, n, generator=gen).cuda()
    x = torch.randn(rows, k, generator=gen).cuda()
    for _ in range(2):
        gemm.linear_tf32x3_wgrad(g, x)
torch.cuda.synchronize()
shapes = torch.tensor(cape_b200.synthetic.CAPE_PYRAMID).cuda()
starts = cape_b200.level_start_index_from_shapes(shapes.cpu()).cuda()
b, m = 128, 8
value = torch.randn(b, 5440, m, 32, generator=gen).cuda()
ref = torch.rand(b, 1, 4, 2, generator=gen).cuda()
off = (torch.randn(b, 1, m, 4, 4, 2, generator=gen) * 3).cuda()
logits = torch.randn(b, 1, m, 16, generator=gen).cuda()
wt = torch.randn(256, 256, generator=gen).cuda()
bias, gamma, beta = torch.zeros(256).cuda(), torch.ones(256).cuda(), torch.zeros(256).cuda()
res = torch.randn(b, 256, generator=gen).cuda()
for _ in range(2):
    K.msda_output_proj(value, shapes, starts, ref, off, logits, wt, bias, residual=res, gamma=gamma, beta=beta)
torch.cuda.synchronize()
print("ok")
