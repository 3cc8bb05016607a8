"""Development aid: %globaltimer stamps of the CTA-pair GEMM per work item (epilogue warp 2 of every CTA and the MMA issuer of
every leader) at the decoder FFN conv shape.  usage: gemm_pair_timing.py [B T Cin N taps]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
import torch  # noqa: E402
from ctts_b200 import capi, engine  # noqa: E402

B, T, Cin, N, taps = [int(v) for v in sys.argv[1:6]] if len(sys.argv) > 5 else (16, 800, 256, 1024, 9)
dev = "cuda:0"
x = torch.randn(B, T, Cin, device=dev)
w = torch.randn(N, taps * Cin, device=dev) / math.sqrt(Cin * taps)
xp, wp = engine.split_planes(x, 2), engine.split_planes(w, 2)
lens = torch.full((B,), T, device=dev, dtype=torch.int64)
kw = dict(taps=taps, act=engine.ACT_RELU, lens=lens, want_fp32=False, want_planes=True)
for _ in range(3):
    engine.gemm_tc(xp, wp, **kw)
torch.cuda.synchronize()
buf = torch.zeros(148 * 64, dtype=torch.int64, device=dev)
capi.call("ctts_debug_set_timing_buffer", buf)
engine.gemm_tc(xp, wp, **kw)
torch.cuda.synchronize()
capi.call("ctts_debug_set_timing_buffer", None)
d = buf.view(148, 64).cpu()
t0 = int(d[:, 0][d[:, 0] > 0].min())
ends = []
for cta in range(148):
    row = d[cta]
    items = []
    for i in range(8):
        b, meta, seen, e = [int(v) for v in row[i * 4:i * 4 + 4]]
        if b == 0:
            break
        items.append((b - t0, seen - t0 if seen else 0, e - t0, meta & 0xFFFF, (meta >> 16) & 0xFFFF, ((meta >> 32) & 0xFF) - 1, (meta >> 40) & 0xFF))
    mma = []
    for i in range(8):
        a, b2, c = [int(v) for v in row[32 + i * 3:32 + i * 3 + 3]]
        if a == 0:
            break
        mma.append((a - t0, b2 - t0, c - t0))
    if items:
        ends.append(items[-1][2])
    if cta in (0, 1, 2, 50, 51, 100, 146, 147):
        print("CTA %3d epilogue items (begin, contributors seen, end us | kb0-kb1 part parts): %s" % (
            cta, "  ".join("%.1f/%.1f/%.1f|%d-%d p%d n%d" % (a / 1e3, s_ / 1e3, e / 1e3, k0, k1, pt, ps) for a, s_, e, k0, k1, pt, ps in items)))
        if mma:
            print("        MMA items (begin, accumulator free, issued us): %s" % "  ".join("%.1f/%.1f/%.1f" % (a / 1e3, b2 / 1e3, c / 1e3) for a, b2, c in mma))
ends.sort()
print("last epilogue end per CTA: min %.1f median %.1f max %.1f us  [%s]" % (ends[0] / 1e3, ends[len(ends) // 2] / 1e3, ends[-1] / 1e3,
      " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("CTTS_"))))
