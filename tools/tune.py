"""Sweep launch geometry of the fast kernels on the bench workload (cape_set_tuning: FWD/BWD_THREADS, FWD/BWD_QPC).
Development tool; prints one line per configuration."""
import ctypes
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import _lib

N, LQ = int(os.environ.get("TUNE_N", 20)), int(os.environ.get("TUNE_LQ", 5440))
dist = os.environ.get("TUNE_DIST", "encoder")
dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[os.environ.get("TUNE_DTYPE", "f32")]
lib = _lib.load()
inp = cape_b200.synthetic.make_inputs(N, LQ, dist=dist, seed=0, device="cuda", dtype=torch.float32)
value = inp["value"].to(dtype)
gout = inp["grad_output"].to(dtype)
loc, attn = inp["sampling_locations"], inp["attention_weights"]
shapes, starts = inp["spatial_shapes"], inp["level_start_index"]
out = torch.empty(N, LQ, 256, device="cuda", dtype=dtype)
gvalue = torch.empty(value.shape, device="cuda")
gloc, gattn = torch.empty_like(loc), torch.empty_like(attn)
dims = _lib.Dims(N, 5440, 8, 32, LQ, 4, 4)
p = lambda t: ctypes.c_void_p(t.data_ptr())
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
vd = 0 if dtype == torch.float32 else 1


def fwd():
    _lib.check(lib.cape_msda_forward(p(value), p(shapes), p(starts), p(loc), p(attn), p(out), ctypes.byref(dims), vd, 0, sp), "f")


def bwd():
    _lib.check(lib.cape_msda_backward(p(gout), p(value), p(shapes), p(starts), p(loc), p(attn), p(gvalue), p(gloc),
                                      p(gattn), ctypes.byref(dims), vd, 0, 0, sp), "b")


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
if which in ("fwd", "both"):
    for th, qpc in itertools.product((128, 256, 512), (128, 256, 512, 1024, 2048)):
        _lib.set_tuning("FWD_THREADS", th), _lib.set_tuning("FWD_QPC", qpc)
        print(f"fwd threads={th:4d} qpc={qpc:4d}  {timeit(fwd):8.1f} us", flush=True)
if which in ("bwd", "both"):
    ths = [int(v) for v in os.environ.get("TUNE_THREADS", "128,256,512").split(",")]
    qpcs = [int(v) for v in os.environ.get("TUNE_QPCS", "64,128,256,512,1024").split(",")]
    for th, qpc in itertools.product(ths, qpcs):
        _lib.set_tuning("BWD_THREADS", th), _lib.set_tuning("BWD_QPC", qpc)
        gvalue.zero_()
        print(f"bwd threads={th:4d} qpc={qpc:4d}  {timeit(bwd):8.1f} us", flush=True)
