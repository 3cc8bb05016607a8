"""Development check of the CTA-pair GEMM (cta_group::2): bit-equality with the single-CTA persistent kernel (same MMA
order per accumulator element), closeness to fp64, and launch time of the decoder FFN conv shape.
Run under `timeout`: a protocol error in a 2-CTA kernel shows up as a hang.   python tools/pair_check.py [swap]"""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "comprehensive-transformer-tts_b200"))
from ctts_b200 import engine  # noqa: E402

DEV = "cuda:0"
if len(sys.argv) > 1 and sys.argv[1] == "swap":
    os.environ["CTTS_PAIR_SWAP_B"] = "1"
os.environ["CTTS_PAIR_VERBOSE"] = "1"


def run(pair, *a, **k):
    os.environ["CTTS_PAIR_GEMM"] = "2" if pair else "0"
    y, yp = engine.gemm_tc(*a, **k)
    torch.cuda.synchronize()
    return y, yp


def case(B, T, Cin, N, taps, act, seed=0):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, Cin, generator=gen)
    w = torch.randn(N, Cin, taps, generator=gen) / math.sqrt(Cin * taps)
    bias = torch.randn(N, generator=gen)
    res = torch.randn(B, T, N, generator=gen)
    lens = torch.tensor([max(T - 9 * b, 1) for b in range(B)])
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), padding=taps // 2).transpose(1, 2)
    ref = {"none": lambda v: v, "gelu": F.gelu, "tanh": torch.tanh, "relu": F.relu}[act](ref) + res.double()
    ref = (ref * (torch.arange(T)[None, :] < lens[:, None]).double()[:, :, None]).float()
    packed = w.permute(0, 2, 1).reshape(N, -1).contiguous().to(DEV)
    xp, wp = engine.split_planes(x.to(DEV)), engine.split_planes(packed)
    kw = dict(act=engine._ACTS[act], residual=res.to(DEV), lens=lens.to(DEV), taps=taps, want_planes=True)
    y0, p0 = run(False, xp, wp, bias.to(DEV), **kw)
    y1, p1 = run(True, xp, wp, bias.to(DEV), **kw)
    e0 = (y0.cpu() - ref).abs().max().item()
    e1 = (y1.cpu() - ref).abs().max().item()
    same = torch.equal(y0, y1) and all(torch.equal(a, b) for a, b in zip(p0.p, p1.p))
    print(f"B{B} T{T} Cin{Cin} N{N} k{taps} {act}: err single {e0:.2e} pair {e1:.2e} bit-equal {same}", flush=True)
    return same and e1 < 2e-4


def timeit(pair, B=16, T=800, Cin=256, N=1024, taps=9, iters=10):
    """(eager us / launch, CUDA-graph us / launch): the graph number is free of host effects."""
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, Cin, generator=gen).to(DEV)
    w = (torch.randn(N, taps * Cin, generator=gen) / math.sqrt(Cin * taps)).to(DEV)
    xp, wp = engine.split_planes(x), engine.split_planes(w)
    os.environ["CTTS_PAIR_GEMM"] = "1" if pair else "0"

    def body():
        for _ in range(iters):
            engine.gemm_tc(xp, wp, act=engine._ACTS["gelu"], taps=taps, want_fp32=False, want_planes=True)

    body()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    body()
    e1.record()
    host = (time.perf_counter() - t0) / iters * 1e6
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / iters * 1e3
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
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return eager, e0.elapsed_time(e1) / (3 * iters) * 1e3, host


if __name__ == "__main__":
    ok = True
    for c in [(2, 256, 256, 512, 1, "none"), (1, 70, 256, 80, 1, "none"), (2, 129, 512, 80, 5, "none"),
              (2, 261, 1024, 256, 1, "none"), (3, 192, 256, 1024, 9, "gelu"), (1, 128, 256, 128, 1, "none"), (3, 300, 256, 1024, 9, "gelu"), (2, 100, 256, 768, 1, "none"),
              (3, 96, 256, 512, 3, "relu"), (5, 160, 512, 512, 5, "tanh"), (16, 800, 256, 1024, 9, "gelu"),
              (16, 800, 512, 512, 5, "tanh"), (16, 800, 256, 768, 1, "none")]:
        ok = case(*c) and ok
    print("ALL OK" if ok else "MISMATCH", flush=True)
    for shape in [dict(), dict(Cin=512, N=512, taps=5), dict(Cin=256, N=768, taps=1)]:
        print(shape, "single eager %.1f graph %.1f host %.1f us | pair eager %.1f graph %.1f host %.1f us"
              % (timeit(False, **shape) + timeit(True, **shape)), flush=True)
