#!/usr/bin/env python
"""Development aid (GPU box): CUDA-event time of each captured stage of the headline inference step (stage A: encoder,
duration model, LengthRegulator scan; stage B: upsampling, pitch / energy, decoder, mel head) and of the host gap between."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import ctts_b200  # noqa: E402
from ctts_b200 import engine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "fs2"
    cfgs, sd, batch, frames = bench.build_workload(seed=0, name=name)
    dev = torch.device("cuda:0")
    net = ctts_b200.CompTransTTS(*cfgs).eval()
    net.load_state_dict(sd, strict=True)
    net.to(dev)
    dev_in = (batch["speakers"].to(dev), batch["texts"].to(dev), batch["src_lens"].to(dev), batch["max_src_len"])
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    events = []
    orig = engine.GraphCache.run

    def run(self, key, fn, tensor_inputs, table=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(self, key, fn, tensor_inputs, table)
        e1.record()
        events.append((e0, e1))
        return r

    engine.GraphCache.run = run
    for _ in range(4):
        net(*dev_in)
    torch.cuda.synchronize()
    rows = []
    for _ in range(20):
        flush.zero_()
        del events[:]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        net(*dev_in)
        s1.record()
        torch.cuda.synchronize()
        ev = list(events)
        rows.append([s0.elapsed_time(s1)] + [a.elapsed_time(b) for a, b in ev] +
                    [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)] +
                    [s0.elapsed_time(ev[0][0]), ev[-1][1].elapsed_time(s1)])
    n = len(rows[0])
    med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(n)]
    k = (n - 3 + 1) // 2
    print("%s: step %.3f ms | stages %s ms | gaps between stages %s ms | before the first stage %.3f ms, after the last %.3f ms" %
          (name, med[0], ", ".join("%.3f" % v for v in med[1:1 + k]), ", ".join("%.3f" % v for v in med[1 + k:-2]),
           med[-2], med[-1]))


if __name__ == "__main__":
    main()
