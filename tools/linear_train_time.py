"""fwd + bwd time of one Linear at the training shapes, nn.Linear vs the 3xTF32 autograd form.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import gemm

dev = "cuda"
M = 20 * 5440


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (k, n) in ((256, 256), (256, 1024), (1024, 256), (256, 128)):
    lin = torch.nn.Linear(k, n).to(dev)
    x = torch.randn(M, k, device=dev, requires_grad=True)
    g = torch.randn(M, n, device=dev)
    out = {}
    for mode in ("fp32", "tf32x3"):
        cape_b200.set_linear_mode(mode)
        fwd = lambda: gemm.linear(lin, x)
        def both():
            y = gemm.linear(lin, x)
            torch.autograd.grad(y, (x, lin.weight, lin.bias), g)
        out[mode] = (t(fwd), t(both))
    cape_b200.set_linear_mode("fp32")
    print(f"{k}->{n}: fwd {out['fp32'][0]:.2f} / {out['tf32x3'][0]:.2f} ms   fwd+bwd {out['fp32'][1]:.2f} / {out['tf32x3'][1]:.2f} ms  (fp32 / tf32x3)")
