"""Weight gradient g^T x of the 3xTF32 linear: MN-major operands read in place (default) vs the earlier transposed copies
(WGRAD_TRANSPOSE=1), accuracy against fp64 and time at the training shapes.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import _lib, gemm


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for rows, n, k in ((1024, 128, 256), (4096, 256, 256), (20 * 5440, 256, 256), (20 * 5440, 1024, 256), (20 * 5440, 256, 1024), (20 * 200, 256, 256)):
    gen = torch.Generator().manual_seed(rows + n + k)
    g = torch.randn(rows, n, generator=gen).cuda()
    x = torch.randn(rows, k, generator=gen).cuda()
    ref = g.double().t() @ x.double()
    line = f"rows {rows:7d} N {n:5d} K {k:5d}:"
    for mode in (0, 1):
        _lib.set_tuning("WGRAD_TRANSPOSE", mode)
        got = gemm.linear_tf32x3_wgrad(g, x)
        torch.cuda.synchronize()
        err = float((got.double() - ref).abs().max() / ref.abs().max())
        line += f"  {'transposed' if mode else 'in place  '} err {err:.2e} {timeit(lambda: gemm.linear_tf32x3_wgrad(g, x)):8.1f} us"
    _lib.set_tuning("WGRAD_TRANSPOSE", 0)
    line += f"  cuBLAS fp32 {timeit(lambda: g.t() @ x):8.1f} us"
    print(line, flush=True)
