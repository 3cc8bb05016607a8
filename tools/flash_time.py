"""Launch time of the fused attention kernel alone (CUDA-graph replay, bench shape: B 16, T 800, lens 800..560)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "comprehensive-transformer-tts_b200"))
from ctts_b200 import engine  # noqa: E402

DEV = "cuda:0"
B, T, C, H = 16, 800, 256, 2
gen = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3 * C, generator=gen).to(DEV)
lens = torch.tensor([8 * (100 - 2 * b) for b in range(B)]).to(DEV)
qp = engine.split_planes(qkv)
iters = 10


def body():
    for _ in range(iters):
        engine.attention_flash(qp, lens, H)


body()
torch.cuda.synchronize()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    body()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        body()
torch.cuda.synchronize()
graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    graph.replay()
e1.record()
torch.cuda.synchronize()
print("attention_flash (V^T planes + fused kernel): %.1f us per call" % (e0.elapsed_time(e1) / (5 * iters) * 1e3))
