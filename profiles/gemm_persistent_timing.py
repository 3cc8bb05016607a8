"""Development aid: %globaltimer stamps of the persistent GEMM per CTA at the decoder's short-K shapes (run on the GPU box).
Prints, per shape: kernel time by CUDA events, time from the first CTA start to the last CTA end, per-CTA time to the first
accumulator (set-up + pipeline fill + first mainloop), per-CTA total, tiles per CTA."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
import torch  # noqa: E402
from ctts_b200 import capi, engine  # noqa: E402

DEV = "cuda:0"


def run(B, T, Cin, N, taps, residual, planes_out, label):
    x = torch.randn(B, T, Cin, device=DEV)
    w = torch.randn(N, taps * Cin, device=DEV) / math.sqrt(Cin * taps)
    xp, wp = engine.split_planes(x, 2), engine.split_planes(w, 2)
    res = torch.randn(B, T, N, device=DEV) if residual else None
    lens = torch.full((B,), T, device=DEV, dtype=torch.int64)
    buf = torch.zeros(4 * 256, dtype=torch.int64, device=DEV)
    kw = dict(taps=taps, residual=res, lens=lens if residual else None, want_fp32=not planes_out, want_planes=planes_out)
    for _ in range(3):
        engine.gemm_tc(xp, wp, **kw)
    torch.cuda.synchronize()
    capi.call("ctts_debug_set_timing_buffer", buf)
    # two back-to-back launches: the second one shows what a kernel costs behind a predecessor of its own kind
    engine.gemm_tc(xp, wp, **kw)
    engine.gemm_tc(xp, wp, **kw)
    torch.cuda.synchronize()
    capi.call("ctts_debug_set_timing_buffer", None)
    d = buf.view(-1, 4).cpu()
    d = d[d[:, 0] != 0]
    if d.shape[0] == 0:
        print("%-26s (not the persistent kernel)" % label)
        return
    t0 = d[:, 0].min()
    tiles = (d[:, 3] & 0xFFFF).float()
    setup = (d[:, 3] >> 16).float()
    print("%-26s first entry -> last end %.1f us | per CTA: entry spread %.1f us, set-up (barriers, TMEM alloc) %.2f us (max %.2f), "
          "entry -> first accumulator %.1f us, total %.1f us (max %.1f), tiles %.1f" % (
              label, (d[:, 2].max() - t0).item() / 1e3, (d[:, 0].max() - t0).item() / 1e3, setup.mean() / 1e3, setup.max() / 1e3,
              (d[:, 1] - d[:, 0]).float().mean() / 1e3, (d[:, 2] - d[:, 0]).float().mean() / 1e3,
              (d[:, 2] - d[:, 0]).max().item() / 1e3, tiles.mean()))


if __name__ == "__main__":
    run(16, 800, 256, 256, 1, True, False, "out-proj K256 N256 +res")
    run(16, 800, 1024, 256, 1, True, False, "FFN2 K1024 N256 +res")
    run(16, 800, 256, 768, 1, False, True, "QKV K256 N768 planes")
    run(16, 800, 512, 512, 5, False, True, "PostNet k5 K2560 N512")
