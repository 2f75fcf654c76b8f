import sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
import cape_b200
from cape_b200 import _lib
lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())
dev = "cuda"
M, N, K = 20 * 5440, 256, 256
x = torch.randn(M, K, device=dev); y = torch.empty(M, N, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
junk = []
for i in range(10):
    junk.append(torch.empty(1000003 * (i * 7 % 5 + 1), device=dev))
    w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
    w_lo = torch.empty_like(w)
    lib.cape_tf32_split_lo(p(w), p(w_lo), w.numel(), sp)
    fn = lambda: lib.cape_linear_tf32x3(p(x), p(w), p(w_lo), p(b), p(y), M, N, K, 0, sp)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"w @ {w.data_ptr():#x} w_lo @ {w_lo.data_ptr():#x}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
