"""Small end-to-end case for compute-sanitizer (one tool per gpurun call): forward, backward, fused op, decode,
generic-dims kernels, host-buffer round trip — all at tiny sizes.  Development tool."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import _lib, synthetic

torch.manual_seed(0)
for shapes, kw in ((synthetic.CAPE_PYRAMID, {}), (((7, 5), (4, 3)), dict(n_heads=3, head_dim=16, n_points=3)),
                   (((9, 5), (4, 4), (2, 3)), {})):
    for dt in (torch.float32, torch.bfloat16):
        inp = synthetic.make_inputs(2, 37, shapes, dist="uniform", seed=1, device="cuda", **kw)
        v = inp["value"].to(dt).requires_grad_(True)
        loc = inp["sampling_locations"].requires_grad_(True)
        attn = inp["attention_weights"].requires_grad_(True)
        out = cape_b200.ms_deform_attn(v, inp["spatial_shapes"], inp["level_start_index"], loc, attn)
        out.backward(inp["grad_output"].to(dt))
n, lq, m, l, p = 3, 5, 8, 4, 4
shapes = torch.tensor(synthetic.CAPE_PYRAMID, device="cuda")
starts = cape_b200.level_start_index_from_shapes(shapes)
value = torch.randn(n, 5440, m, 32, device="cuda", requires_grad=True)
ref = torch.rand(n, lq, l, 2, device="cuda", requires_grad=True)
off = torch.randn(n, lq, m, l, p, 2, device="cuda", requires_grad=True)
logits = torch.randn(n, lq, m, l * p, device="cuda", requires_grad=True)
cape_b200.ms_deform_attn_fused(value, shapes, starts, ref, off, logits).sum().backward()
with torch.no_grad():
    cape_b200.ms_deform_attn_decode(value[:, :, :, :], shapes, starts, ref[:, :1], off[:, :1], logits[:, :1])
lib = _lib.load()
inp = synthetic.make_inputs(2, 64, dist="encoder", seed=2)
dims = _lib.Dims(2, 5440, 8, 32, 64, 4, 4)
ws_bytes = lib.cape_msda_host_workspace_bytes(ctypes.byref(dims), 1)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
pin = {k: inp[k].contiguous().pin_memory() for k in ("value", "sampling_locations", "attention_weights", "grad_output")}
res = [torch.empty(s).pin_memory() for s in ((2, 64, 256), (2, 5440, 8, 32), (2, 64, 8, 4, 4, 2), (2, 64, 8, 4, 4))]
q = lambda t: ctypes.c_void_p(t.data_ptr())
_lib.check(lib.cape_msda_forward_backward_host(q(pin["value"]), q(inp["spatial_shapes"]), q(inp["level_start_index"]),
                                               q(pin["sampling_locations"]), q(pin["attention_weights"]),
                                               q(pin["grad_output"]), q(res[0]), q(res[1]), q(res[2]), q(res[3]),
                                               ctypes.byref(dims), q(ws), ws_bytes,
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "host")
torch.cuda.synchronize()
print("sanitize_case ok, launches:", cape_b200.launch_count())
