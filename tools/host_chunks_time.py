"""e2e time of cape_msda_forward_backward_host (pinned host buffers, bench shape) per HOST_CHUNKS setting.  Development tool."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import torch
import cape_b200
from cape_b200 import _lib

lib = _lib.load()
N, LQ = 20, 5440
inp = cape_b200.synthetic.make_inputs(N, LQ, dist="encoder", seed=0)
pin = {k: v.pin_memory() for k, v in inp.items()}
out = torch.empty(N, LQ, 256).pin_memory()
gv = torch.empty_like(inp["value"]).pin_memory()
gl = torch.empty_like(inp["sampling_locations"]).pin_memory()
ga = torch.empty_like(inp["attention_weights"]).pin_memory()
dims = _lib.Dims(N, 5440, 8, 32, LQ, 4, 4)
need = lib.cape_msda_host_workspace_bytes(ctypes.byref(dims), 1)
ws = torch.empty(need, dtype=torch.uint8, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: _lib.check(lib.cape_msda_forward_backward_host(
    p(pin["value"]), p(pin["spatial_shapes"]), p(pin["level_start_index"]), p(pin["sampling_locations"]),
    p(pin["attention_weights"]), p(pin["grad_output"]), p(out), p(gv), p(gl), p(ga), ctypes.byref(dims), p(ws),
    ctypes.c_size_t(need), sp), "host")
ref = None
for chunks in (4, 8, 10, 20, 32):
    _lib.set_tuning("HOST_CHUNKS", chunks)
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        call()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 8 * 1e3
    chk = float(gv.double().sum() + out.double().sum())
    ref = chk if ref is None else ref
    print(f"HOST_CHUNKS={chunks:3d}  {ms:7.3f} ms/step  {1058.4 / ms:7.1f} GB/s  checksum {'same' if abs(chk - ref) <= 1e-6 * abs(ref) else 'DIFFERENT'}", flush=True)
