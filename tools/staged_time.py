"""Time the L1 kernels against the shared-memory staged kernels on the bench workload (N=20, Lq=S=5440), forward and
backward, per shared-memory budget.  Development tool; prints one line per configuration."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import _lib

N, LQ = int(os.environ.get("TUNE_N", 20)), int(os.environ.get("TUNE_LQ", 5440))
lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=20):
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


which = sys.argv[1] if len(sys.argv) > 1 else "both"
for dist in ("encoder", "uniform"):
    inp = cape_b200.synthetic.make_inputs(N, LQ, dist=dist, seed=0, device="cuda")
    loc, attn = inp["sampling_locations"], inp["attention_weights"]
    shapes, starts = inp["spatial_shapes"], inp["level_start_index"]
    dims = _lib.Dims(N, 5440, 8, 32, LQ, 4, 4)
    for name, dt, code in (("f32", torch.float32, 0), ("bf16", torch.bfloat16, 1)):
        value, gout = inp["value"].to(dt), inp["grad_output"].to(dt)
        out = torch.empty(N, LQ, 256, device="cuda", dtype=dt)
        gvalue = torch.empty(inp["value"].shape, device="cuda")
        gloc, gattn = torch.empty_like(loc), torch.empty_like(attn)
        fwd = lambda: _lib.check(lib.cape_msda_forward(p(value), p(shapes), p(starts), p(loc), p(attn), p(out),
                                                       ctypes.byref(dims), code, 0, sp), "f")
        bwd = lambda: _lib.check(lib.cape_msda_backward(p(gout), p(value), p(shapes), p(starts), p(loc), p(attn), p(gvalue),
                                                        p(gloc), p(gattn), ctypes.byref(dims), code, 0, 1, sp), "b")
        if which in ("fwd", "both"):
            _lib.set_tuning("FWD_STAGED", 2)
            print(f"{dist:8s} {name:4s} fwd L1 kernel            {timeit(fwd):8.1f} us", flush=True)
            _lib.set_tuning("FWD_STAGED", 1)
            for kb in (8, 48, 200):
                _lib.set_tuning("FWD_STAGED_KB", kb)
                print(f"{dist:8s} {name:4s} fwd staged {kb:3d} KB        {timeit(fwd):8.1f} us", flush=True)
            _lib.set_tuning("FWD_STAGED_KB", 0)
        if which in ("bwd", "both"):
            for mode in (int(m) for m in os.environ.get("TUNE_BWD_MODES", "1,2").split(",")):
                _lib.set_tuning("BWD_MODE", mode)
                for prof in (int(x) for x in os.environ.get("TUNE_PROFILE", "0").split(",")) if mode == 5 else (0,):
                    _lib.set_tuning("PROFILE", prof)
                    for qpc in (int(x) for x in os.environ.get("TUNE_QPC", "0").split(",")) if mode == 5 else (0,):
                        _lib.set_tuning("BWD_QPC", qpc)
                        print(f"{dist:8s} {name:4s} bwd mode {mode} prof {prof} qpc {qpc:4d} (memset incl.) {timeit(bwd):8.1f} us", flush=True)
                _lib.set_tuning("PROFILE", 0)
                _lib.set_tuning("BWD_QPC", 0)
            _lib.set_tuning("BWD_MODE", 0)
