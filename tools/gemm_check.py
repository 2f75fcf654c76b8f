import sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
import cape_b200
from cape_b200 import _lib
lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())
torch.manual_seed(0)
dev = "cuda"
for (M, N, K) in [(128, 128, 32), (128, 128, 256), (300, 256, 256), (4096, 1024, 256), (1000, 256, 1024)]:
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    w_lo = torch.empty_like(w)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.cape_tf32_split_lo(p(w), p(w_lo), w.numel(), sp), "split")
    y = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.cape_linear_tf32x3(p(x), p(w), p(w_lo), p(b), p(y), M, N, K, 0, sp), "gemm")
    torch.cuda.synchronize()
    ref64 = (x.double() @ w.double().t() + b.double())
    ref32 = torch.nn.functional.linear(x, w, b)
    e = lambda a: float((a.double() - ref64).abs().max() / ref64.abs().max())
    print(M, N, K, "err tf32x3 %.2e  err torch fp32 %.2e  nan=%d" % (e(y), e(ref32), int(torch.isnan(y).sum())))
# timing
M, N, K = 128 * 5440, 256, 256
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / 16; b = torch.randn(N, device=dev)
w_lo = torch.empty_like(w); y = torch.empty(M, N, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.cape_tf32_split_lo(p(w), p(w_lo), w.numel(), sp)
for name, fn in (("tf32x3", lambda: lib.cape_linear_tf32x3(p(x), p(w), p(w_lo), p(b), p(y), M, N, K, 0, sp)),
                 ("torch fp32", lambda: torch.nn.functional.linear(x, w, b))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")

# the other encoder shapes (FFN) at N = 128 x 5440 rows
for (M, N, K, relu) in [(128 * 5440, 1024, 256, 1), (128 * 5440, 256, 1024, 0), (128 * 5440, 128, 256, 0)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    w_lo = torch.empty_like(w); y = torch.empty(M, N, device=dev)
    lib.cape_tf32_split_lo(p(w), p(w_lo), w.numel(), sp)
    for name, fn in (("tf32x3", lambda: lib.cape_linear_tf32x3(p(x), p(w), p(w_lo), p(b), p(y), M, N, K, relu, sp)),
                     ("torch fp32", lambda: torch.nn.functional.linear(x, w, b))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"M={M} N={N} K={K} {name}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")
    del x, y
