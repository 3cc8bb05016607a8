#!/usr/bin/env python
"""Development aid (GPU box): CUDA-event time of one planes-in / planes-out Conv1d GEMM.  usage: conv_gemm_time.py B T Cin N taps"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
import torch  # noqa: E402
from ctts_b200 import engine  # noqa: E402

B, T, Cin, N, taps = [int(v) for v in sys.argv[1:6]]
dev = "cuda:0"
x = torch.randn(B, T, Cin, device=dev)
w = torch.randn(N, taps * Cin, device=dev) / math.sqrt(Cin * taps)
bias = torch.randn(N, device=dev)
xp, wp = engine.split_planes(x, 2), engine.split_planes(w, 2)
lens = torch.full((B,), T, device=dev, dtype=torch.int64)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
kw = dict(taps=taps, bias=bias, act=engine.ACT_RELU, lens=lens, want_fp32=False, want_planes=True)
for _ in range(3):
    engine.gemm_tc(xp, wp, **kw)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    engine.gemm_tc(xp, wp, **kw)
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1000)
ts.sort()
flop = 2.0 * B * T * Cin * taps * N
print("B %d T %d Cin %d N %d taps %d: median %.1f us (min %.1f)  %.0f TFLOP/s algorithmic  [%s]" %
      (B, T, Cin, N, taps, ts[10], ts[0], flop / ts[10] * 1e-6, " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("CTTS_"))))
