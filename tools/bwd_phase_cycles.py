"""Where the MMA warp of the staged backward kernel (BWD_MODE 2) spends its cycles: per-phase clock64() totals of CTA 0
(cape_debug_counters).  Development tool."""
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
gv = torch.empty_like(inp["value"])
gl, ga = torch.empty_like(inp["sampling_locations"]), torch.empty_like(inp["attention_weights"])
dims = _lib.Dims(N, 5440, 8, 32, LQ, 4, 4)
bwd = lambda: _lib.check(lib.cape_msda_backward(p(inp["grad_output"]), p(inp["value"]), p(inp["spatial_shapes"]),
                                                p(inp["level_start_index"]), p(inp["sampling_locations"]),
                                                p(inp["attention_weights"]), p(gv), p(gl), p(ga), ctypes.byref(dims), 0, 0, 1, sp), "b")
_lib.set_tuning("BWD_MODE", int(os.environ.get("MODE", 2)))
_lib.set_tuning("PROFILE", 1)
bwd()
torch.cuda.synchronize()
out = (ctypes.c_longlong * 16)()
_lib.check(lib.cape_debug_counters(out, 1), "dbg")
bwd()
torch.cuda.synchronize()
_lib.check(lib.cape_debug_counters(out, 1), "dbg")
names = ["loop top", "wait prev MMAs", "-", "-", "wait 31 columns", "MMA issue", "tail wait", "-", "-", "batches", "-"]
batches = max(1, out[9])
for i, n in enumerate(names):
    print(f"{n:20s} {out[i]:12d}   per batch {out[i] / batches:10.1f}")
