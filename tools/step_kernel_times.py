"""Warm per-launch time of each decode-step kernel (100 back-to-back launches captured in one CUDA graph, replayed).
Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200
from cape_b200 import decode_ops as K

dev = torch.device("cuda", 0)
B = int(os.environ.get("GEN_N", "128"))
torch.manual_seed(0)
x = torch.randn(B, 256, device=dev)
x2 = torch.randn(B, 256, device=dev)
h = torch.randn(B, 1024, device=dev)
w256 = torch.randn(256, 256, device=dev) * 0.05
w768 = torch.randn(256, 768, device=dev) * 0.05
w384 = torch.randn(256, 384, device=dev) * 0.05
w1024 = torch.randn(256, 1024, device=dev) * 0.05
w1024t = torch.randn(1024, 256, device=dev) * 0.05
b256, b768, b384, b1024 = (torch.randn(n, device=dev) for n in (256, 768, 384, 1024))
g, be = torch.ones(256, device=dev), torch.zeros(256, device=dev)
ref = torch.rand(B, 2, device=dev)
valid = torch.ones(B, 4, 2, device=dev)
w3, b3 = torch.randn(2, 256, device=dev) * 0.05, torch.zeros(2, device=dev)
wc, bc = torch.randn(3, 256, device=dev), torch.zeros(3, device=dev)
dim_t = 10000 ** (2 * (torch.arange(128, device=dev, dtype=torch.float32) // 2) / 128)
kc, vc = torch.randn(B, 101, 256, device=dev), torch.randn(B, 101, 256, device=dev)
pos = torch.full((1,), 50, dtype=torch.int64, device=dev)
qkv = torch.randn(B, 768, device=dev)
sk, sv = torch.randn(B, 17, 256, device=dev), torch.randn(B, 17, 256, device=dev)
bias = torch.zeros(B, 17, device=dev)
value = torch.randn(B, 5440, 8, 32, device=dev)
shapes = torch.tensor(cape_b200.synthetic.CAPE_PYRAMID, device=dev)
starts = cape_b200.level_start_index_from_shapes(shapes)
off = torch.randn(B, 1, 8, 4, 4, 2, device=dev)
logits = torch.randn(B, 1, 8, 16, device=dev)
refl = torch.rand(B, 1, 4, 2, device=dev)

cases = {
    "skinny 256->256 bias": lambda: K.skinny_linear(x, w256, b256),
    "skinny 256->256 +x2": lambda: K.skinny_linear(x, w256, b256, x2=x2),
    "skinny 256->256 res+LN": lambda: K.skinny_linear(x, w256, b256, residual=x2, gamma=g, beta=be),
    "skinny sine->256 LN": lambda: K.skinny_linear(ref, w256, b256, gamma=g, beta=be, sine_dim_t=dim_t),
    "skinny 256->768": lambda: K.skinny_linear(x, w768, b768),
    "skinny 256->384 split": lambda: K.skinny_linear_split(x, w384, b384, 256, x2=x2),
    "skinny 256->1024 relu": lambda: K.skinny_linear(x, w1024, b1024, relu=True),
    "skinny 1024->256 res+LN": lambda: K.skinny_linear(h, w1024t, b256, residual=x2, gamma=g, beta=be),
    "coord head refine": lambda: K.coord_head_refine(x, w256, b256, w3, b3, ref, valid),
    "tiny 256->3": lambda: K.tiny_linear(x, wc, bc),
    "attention self (51 keys)": lambda: K.decode_attention(qkv[:, :256], kc, vc, qkv[:, 256:512], qkv[:, 512:], pos),
    "attention support (17 keys)": lambda: K.decode_attention(x, sk, sv, key_bias=bias),
    "msda decode": lambda: torch.ops.cape.ms_deform_attn_decode(value, shapes, starts, refl, off, logits),
    "msda decode + output_proj + res + LN (one launch)": lambda: K.msda_output_proj(value, shapes, starts, refl, off, logits, w256, b256,
                                                                                   residual=x2, gamma=g, beta=be),
    "torch add (baseline launch)": lambda: torch.add(x, x2),
}
reps = 100
for name, fn in cases.items():
    fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1) / (5 * reps) * 1e3:7.2f} us")
