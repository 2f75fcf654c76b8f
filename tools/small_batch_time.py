"""fwd / bwd time of the default kernels for N in {2, 4, 20} at Lq = 5440 (and a few other Lq), warm L2.  Development tool."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import _lib

lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for n, lq in ((2, 5440), (4, 5440), (20, 5440), (2, 1000), (2, 20000), (20, 200)):
    inp = cape_b200.synthetic.make_inputs(n, lq, dist="encoder", seed=0, device="cuda")
    out = torch.empty(n, lq, 256, device="cuda")
    gv = torch.empty_like(inp["value"])
    gl, ga = torch.empty_like(inp["sampling_locations"]), torch.empty_like(inp["attention_weights"])
    dims = _lib.Dims(n, 5440, 8, 32, lq, 4, 4)
    fwd = lambda: _lib.check(lib.cape_msda_forward(p(inp["value"]), p(inp["spatial_shapes"]), p(inp["level_start_index"]),
                                                   p(inp["sampling_locations"]), p(inp["attention_weights"]), p(out),
                                                   ctypes.byref(dims), 0, 0, sp), "f")
    bwd = lambda: _lib.check(lib.cape_msda_backward(p(inp["grad_output"]), p(inp["value"]), p(inp["spatial_shapes"]),
                                                    p(inp["level_start_index"]), p(inp["sampling_locations"]),
                                                    p(inp["attention_weights"]), p(gv), p(gl), p(ga), ctypes.byref(dims), 0, 0, 1,
                                                    sp), "b")
    res = []
    for label, fq, bq in (("balanced", 0, 0), ("fixed 512/128", 512, 128)):
        _lib.set_tuning("FWD_QPC", fq)
        _lib.set_tuning("BWD_QPC", bq)
        res.append(f"{label}: fwd {timeit(fwd):7.1f} us  bwd {timeit(bwd):7.1f} us")
    print(f"N={n:2d} Lq={lq:5d}   " + "   |   ".join(res), flush=True)
