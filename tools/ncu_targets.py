"""One launch of each sampling-kernel variant at the bench shape (N=20, Lq=S=5440, fp32, encoder-like) for ncu:
forward L1 kernel, forward staged (200 KB), backward L1 kernel (mode 1), backward staged + tensor-core scatter (mode 2),
backward staged with the covered levels' REDs dropped (mode 4, timing floor only), backward small CTAs + tcgen05 scatter of
the coarsest level (mode 5).  Prints CUDA-event times when run plain."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import _lib

lib = _lib.load()
N, LQ = 20, 5440
inp = cape_b200.synthetic.make_inputs(N, LQ, dist="encoder", seed=0, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
out = torch.empty(N, LQ, 256, device="cuda")
gv = torch.empty_like(inp["value"])
gl, ga = torch.empty_like(inp["sampling_locations"]), torch.empty_like(inp["attention_weights"])
dims = _lib.Dims(N, 5440, 8, 32, LQ, 4, 4)
fwd = lambda: _lib.check(lib.cape_msda_forward(p(inp["value"]), p(inp["spatial_shapes"]), p(inp["level_start_index"]),
                                               p(inp["sampling_locations"]), p(inp["attention_weights"]), p(out),
                                               ctypes.byref(dims), 0, 0, sp), "f")
bwd = lambda: _lib.check(lib.cape_msda_backward(p(inp["grad_output"]), p(inp["value"]), p(inp["spatial_shapes"]),
                                                p(inp["level_start_index"]), p(inp["sampling_locations"]),
                                                p(inp["attention_weights"]), p(gv), p(gl), p(ga), ctypes.byref(dims), 0, 0, 0, sp), "b")


def run(label, fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:45s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us", flush=True)


run("forward  L1 kernel (default)", fwd)
_lib.set_tuning("FWD_STAGED", 1)
run("forward  staged, levels 1-3 in shared memory", fwd)
_lib.set_tuning("FWD_STAGED", 0)
gv.zero_()
run("backward L1 kernel + REDs (default, mode 1)", bwd)
for mode, label in ((2, "backward staged + tensor-core scatter (mode 2)"), (4, "backward staged, coarse REDs dropped (mode 4)"),
                    (5, "backward small CTAs + tensor-core scatter of the 8x8 level (mode 5)")):
    _lib.set_tuning("BWD_MODE", mode)
    run(label, bwd)
_lib.set_tuning("BWD_MODE", 0)
