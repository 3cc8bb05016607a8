"""Development aid: per-CTA cycle breakdown of the tensor-core GEMM at the bench shapes (run on the GPU box).

    python profiles/gemm_cta_timing.py            # prints setup / mainloop / epilogue cycles per CTA for a few shapes
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
import torch  # noqa: E402
from ctts_b200 import capi, engine  # noqa: E402

DEV = "cuda:0"


def run(B, T, Cin, N, taps, np_, want_planes, want_fp32, label):
    x = torch.randn(B, T, Cin, device=DEV)
    w = torch.randn(N, taps * Cin, device=DEV) / math.sqrt(Cin * taps)
    xp, wp = engine.split_planes(x, np_), engine.split_planes(w, np_)
    tiles = B * ((T + 127) // 128)
    n_cta = (tiles + 1) * max(1, (N + 127) // 128)
    buf = torch.zeros(4 * n_cta + 64, dtype=torch.int64, device=DEV)
    for _ in range(2):
        engine.gemm_tc(xp, wp, taps=taps, want_fp32=want_fp32, want_planes=want_planes)
    torch.cuda.synchronize()
    capi.call("ctts_debug_set_timing_buffer", buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.gemm_tc(xp, wp, taps=taps, want_fp32=want_fp32, want_planes=want_planes)
    e1.record()
    torch.cuda.synchronize()
    capi.call("ctts_debug_set_timing_buffer", None)
    d = buf[: 4 * n_cta].view(-1, 4).cpu()
    d = d[d[:, 0] != 0]
    setup = (d[:, 1] - d[:, 0]).float()
    main = (d[:, 2] - d[:, 1]).float()
    epi = (d[:, 3] - d[:, 2]).float()
    span = (d[:, 3].max() - d[:, 0].min()).item()
    nkb = taps * ((Cin + 63) // 64)
    print("%-28s ctas %4d  kernel %.1f us | per CTA cycles: setup %6.0f  mainloop %7.0f (%5.0f / k-block)  epilogue %6.0f | "
          "first-start..last-end %d cyc" % (label, d.shape[0], e0.elapsed_time(e1) * 1e3, setup.mean(), main.mean(),
                                            main.mean() / nkb, epi.mean(), span))


if __name__ == "__main__":
    run(16, 800, 256, 1024, 9, 2, True, False, "FFN1 conv k9 (planes out)")
    run(16, 800, 1024, 256, 1, 2, False, True, "FFN2 (fp32 out)")
    run(16, 800, 256, 768, 1, 2, True, False, "QKV (planes out)")
    run(16, 800, 512, 512, 5, 2, True, False, "PostNet conv k5")
    run(16, 100, 256, 1024, 9, 3, True, False, "encoder FFN1 bf16x6")
