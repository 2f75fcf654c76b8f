"""One IncrementalDecoder step (6 layers, N=128) for a kernel launch list under ncu.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cape_b200

dev = torch.device("cuda", 0)
n, tokens = 128, 100
torch.manual_seed(0)
layers = [cape_b200.TransformerDecoderLayer(256, 1024, 0.1, "relu", 4, 8, 4).to(dev).eval() for _ in range(6)]
shapes = torch.tensor(cape_b200.synthetic.CAPE_PYRAMID, device=dev)
starts = cape_b200.level_start_index_from_shapes(shapes)
memory = torch.randn(n, 5440, 256, device=dev)
sup = torch.randn(n, 17, 256, device=dev)
sup_mask = torch.zeros(n, 17, dtype=torch.bool, device=dev)
tgt = torch.randn(n, 1, 256, device=dev)
qpos = torch.randn(n, 1, 256, device=dev)
ref = torch.rand(n, 1, 4, 2, device=dev)
dec = cape_b200.IncrementalDecoder(layers, n, tokens, dev)
dec.reset(memory, shapes, starts, sup, sup_mask)
use_graph = os.environ.get("NO_GRAPH") is None
for i in range(3):
    dec.step(i, tgt, qpos, ref, use_graph=use_graph)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("step")
dec.step(50, tgt, qpos, ref, use_graph=use_graph)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok")
