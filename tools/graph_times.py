#!/usr/bin/env python
"""Development aid (GPU box): in-graph time of the pieces of the headline forward.  Each piece is captured into its own
CUDA graph and replayed 20 times with an L2 flush in between (CUDA events around the replay), so the figures include the
real launch gaps of a graph and warm instruction caches -- unlike an ncu launch list, which serialises cold launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import ctts_b200  # noqa: E402
from ctts_b200 import engine  # noqa: E402
from ctts_b200.engine import gemm_tc, layernorm_planes, attention_tc, split_planes  # noqa: E402

DEV = torch.device("cuda:0")
FLUSH = None


def graph_time(label, fn, reps=20):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000.0)
    ts.sort()
    print("%-58s %8.1f us  (min %.1f)" % (label, ts[len(ts) // 2], ts[0]))
    return ts[len(ts) // 2]


def layer_pieces(prep, P, pre, x, lens, n_head, kernel, act, n, label):
    W = prep.w
    tag = "#planes" if n == 2 else "#planes3"
    lp = "%slayers.0.op." % pre
    st = {}

    def ln1():
        st["hp"] = layernorm_planes(x, P[lp + "layer_norm1.weight"], P[lp + "layer_norm1.bias"], 1e-12, n=n)[1]

    def qkv():
        st["qkv"] = gemm_tc(st["hp"], W[lp + "self_attn.in_proj_weight" + tag], want_fp32=False, want_planes=True)[1]

    def attn():
        st["ap"] = attention_tc(st["qkv"], lens, n_head)

    def outp():
        gemm_tc(st["ap"], W[lp + "self_attn.out_proj.weight" + tag], residual=x, lens=lens, out=x)

    def ffn1():
        st["fp"] = gemm_tc(st["hp"], W[lp + "ffn.ffn_1.weight" + tag], P[lp + "ffn.ffn_1.bias"], alpha=kernel ** -0.5,
                           act=act, taps=kernel, want_fp32=False, want_planes=True)[1]

    def ffn2():
        gemm_tc(st["fp"], W[lp + "ffn.ffn_2.weight" + tag], P[lp + "ffn.ffn_2.bias"], residual=x, lens=lens, out=x)

    def whole():
        ln1(); qkv(); attn(); outp(); ln1(); ffn1(); ffn2()

    tot = 0.0
    for name, f in (("LayerNorm -> planes", ln1), ("QKV projection", qkv), ("attention", attn), ("out-proj + residual", outp),
                    ("FFN conv k%d + act" % kernel, ffn1), ("FFN linear + residual", ffn2)):
        tot += graph_time("%s: %s" % (label, name), f)
    w = graph_time("%s: whole layer (7 pieces in one graph)" % label, whole)
    print("%-58s %8.1f us" % ("%s: sum of pieces (LN counted twice)" % label, tot + 0))
    return w


def main():
    cfgs, sd, batch, frames = bench.build_workload(seed=0, name="fs2")
    net = ctts_b200.CompTransTTS(*cfgs).eval()
    net.load_state_dict(sd, strict=True)
    net.to(DEV)
    out = net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
    prep = net._prepared
    P = prep.params()
    cfg = net.model_config
    c = cfg["transformer_fs2"]
    act = engine._ACTS[cfg["variance_predictor"]["ffn_act"]]
    src_lens = batch["src_lens"].to(DEV)
    mel_lens = out[9].clone()
    B, S = batch["texts"].shape
    M = int(out[1].shape[1])
    texts = batch["texts"].to(DEV)
    xe = torch.randn(B, S, 256, device=DEV)
    xd = torch.randn(B, M, 256, device=DEV)
    print("B %d  S %d  M %d" % (B, S, M))
    graph_time("encoder (embedding + 4 layers + LN)", lambda: engine._encode(net, prep, P, cfg, texts, src_lens))
    graph_time("duration predictor", lambda: engine.duration_predictor(prep, P, cfg, xe, src_lens))
    layer_pieces(prep, P, "encoder.", xe, src_lens, c["encoder_head"], c["ffn_kernel_size"], act, 3, "encoder layer")
    graph_time("decoder (positions + 6 layers + LN) + mel head", lambda: engine._decode(net, prep, P, cfg, xd.clone(), mel_lens))
    layer_pieces(prep, P, "decoder.", xd, mel_lens, c["decoder_head"], c["ffn_kernel_size"], act, 2, "decoder layer")
    dp = split_planes(xd, 2)
    graph_time("mel_linear + PostNet", lambda: engine.mel_head(prep, P, xd, dp))


if __name__ == "__main__":
    main()
