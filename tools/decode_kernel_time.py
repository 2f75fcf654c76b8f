"""Time the fused decode kernel alone (N = 128, Lq = 1) with the point-parallel and the quad kernels.  Dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cape_b200
n = 128
shapes = torch.tensor(cape_b200.synthetic.CAPE_PYRAMID, device="cuda")
starts = cape_b200.level_start_index_from_shapes(shapes)
value = torch.randn(n, 5440, 8, 32, device="cuda")
ref = torch.rand(n, 1, 4, 2, device="cuda")
off = torch.randn(n, 1, 8, 4, 4, 2, device="cuda") * 3
logits = torch.randn(n, 1, 8, 16, device="cuda")
def t(reps=200):
    for _ in range(5): cape_b200.ms_deform_attn_decode(value, shapes, starts, ref, off, logits)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): cape_b200.ms_deform_attn_decode(value, shapes, starts, ref, off, logits)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps // 20): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("point kernel  %.2f us" % t())
os.environ["CAPE_FWD_POINT_MAX_QM"] = "1"
print("quad kernel   %.2f us" % t())
