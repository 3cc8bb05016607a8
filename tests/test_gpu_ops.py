"""GPU: every kernel of libctts_b200, called through the C ABI, against the CPU oracle / plain fp32 torch.

Tolerances: FP32 kernels 2e-5 abs + 1e-4 rel (summation-order noise only); integer / index results bit-exact.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from ctts_b200 import capi, engine  # noqa: E402
from oracle import ctts_oracle as O  # noqa: E402

DEV = "cuda:0"


def g(seed):
    gen = torch.Generator()
    gen.manual_seed(seed)
    return gen


def close(a, b, atol=2e-5, rtol=1e-4, msg=""):
    np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), atol=atol, rtol=rtol, err_msg=msg)


def stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,T,Cin,N,taps,act", [
    (2, 37, 256, 768, 1, "none"), (3, 100, 256, 1024, 9, "gelu"), (2, 130, 1024, 256, 1, "none"),
    (2, 300, 80, 512, 5, "tanh"), (2, 129, 512, 80, 5, "none"), (4, 64, 256, 1, 1, "none"),
    (2, 50, 128, 256, 5, "relu"), (1, 16, 256, 11, 1, "none"), (2, 257, 256, 256, 3, "relu"),
    (1, 16, 256, 256, 1, "relu"), (16, 100, 256, 1, 1, "none"), (2, 800, 256, 11, 1, "none"),   # skinny-linear path
])
def test_conv1d_gemm_fp32(B, T, Cin, N, taps, act):
    x = torch.randn(B, T, Cin, generator=g(1))
    w = torch.randn(N, Cin, taps, generator=g(2)) / math.sqrt(Cin * taps)
    bias = torch.randn(N, generator=g(3))
    res = torch.randn(B, T, N, generator=g(4))
    lens = torch.tensor([max(T - 7 * b, 1) for b in range(B)])
    sc, sh = torch.rand(N, generator=g(5)) + 0.5, torch.randn(N, generator=g(6))
    alpha = 0.37
    ref = F.conv1d(x.transpose(1, 2), w, bias, padding=taps // 2).transpose(1, 2) * alpha
    ref = ref * sc + sh
    ref = {"none": lambda v: v, "gelu": F.gelu, "tanh": torch.tanh, "relu": F.relu}[act](ref) + res
    ref = ref * (torch.arange(T)[None, :] < lens[:, None]).float()[:, :, None]
    wd = w.to(DEV)
    packed = torch.empty(N, taps * Cin, device=DEV)
    capi.call("ctts_pack_conv_weight", wd, N, Cin, taps, packed, stream())
    assert torch.equal(packed.cpu(), w.permute(0, 2, 1).reshape(N, -1))
    y = engine.conv_gemm(x.to(DEV), packed, bias.to(DEV), alpha=alpha, bn=(sc.to(DEV), sh.to(DEV)),
                         act=engine._ACTS[act], residual=res.to(DEV), lens=lens.to(DEV), taps=taps)
    close(y, ref, atol=5e-5)


def test_conv1d_gemm_rejects_bad_shapes():
    x = torch.zeros(1, 8, 24, device=DEV)
    w = torch.zeros(8, 24, device=DEV)
    with pytest.raises(capi.CttsError):
        engine.conv_gemm(x, w)  # Cin % 16 != 0


@pytest.mark.parametrize("rows,C,eps", [(77, 256, 1e-12), (300, 128, 1e-5), (5, 1024, 1e-12), (64, 80, 1e-5)])
def test_layernorm(rows, C, eps):
    x = torch.randn(1, rows, C, generator=g(7)) * 3 + 1
    x[0, 3] = 0  # LN(0) = beta
    gm, bt = torch.randn(C, generator=g(8)), torch.randn(C, generator=g(9))
    lens = torch.tensor([rows - 2])
    ref = F.layer_norm(x, (C,), gm, bt, eps)
    y = engine.layernorm(x.to(DEV), gm.to(DEV), bt.to(DEV), eps)
    close(y, ref)
    assert torch.allclose(y[0, 3].cpu(), bt, atol=1e-6)
    y2 = engine.layernorm(x.to(DEV), gm.to(DEV), bt.to(DEV), eps, lens.to(DEV))
    ref2 = ref.clone()
    ref2[0, rows - 2:] = 0
    close(y2, ref2)


@pytest.mark.parametrize("B,T,C,H", [(3, 100, 256, 2), (2, 333, 256, 2), (2, 70, 256, 8), (1, 33, 128, 2)])
def test_attention(B, T, C, H):
    qkv = torch.randn(B, T, 3 * C, generator=g(10))
    lens = torch.tensor([max(T - 13 * b, 1) for b in range(B)])
    dh = C // H
    q, k, v = qkv.split(C, -1)
    q = q.view(B, T, H, dh).transpose(1, 2) / math.sqrt(dh)
    k = k.view(B, T, H, dh).transpose(1, 2)
    v = v.view(B, T, H, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    pad = torch.arange(T)[None, :] >= lens[:, None]
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, C)
    ref = ref * (~pad).float()[:, :, None]
    out = engine.attention(qkv.to(DEV), lens.to(DEV), H)
    close(out, ref, atol=3e-5)


def test_embed_and_positions():
    B, S, C, V = 3, 50, 256, 361
    tok = torch.randint(1, V, (B, S), generator=g(11))
    lens = torch.tensor([50, 41, 7])
    for b in range(B):
        tok[b, lens[b]:] = 0
    tok[0, 5] = 0  # an in-sequence pad symbol: position must not advance (utils/tools.py:640-652)
    table = torch.randn(V, C, generator=g(12))
    table[0] = 0
    pe = O.sinusoid_table_fs2(2048, C)
    x = torch.empty(B, S, C, device=DEV)
    word = torch.empty(B, S, C, device=DEV)
    capi.call("ctts_embed_tokens", tok.to(DEV), table.to(DEV), pe.to(DEV), 2048, 16.0, B, S, C, V, x, word,
              lens.to(DEV), 0, stream())
    w_ref = 16.0 * F.embedding(tok, table)
    x_ref = (w_ref + O.fs2_positional(tok, C, 1000)) * (torch.arange(S)[None] < lens[:, None]).float()[:, :, None]
    close(word, w_ref, atol=1e-6)
    close(x, x_ref, atol=1e-6)
    # decoder-style positions from x[..., 0] != 0, scaled by a device scalar, then masked
    h = torch.randn(B, S, C, generator=g(13))
    h[1, 4, 0] = 0
    alpha = torch.tensor([0.73])
    ref = (h + alpha * O.fs2_positional(h[..., 0], C, 2000)) * (torch.arange(S)[None] < lens[:, None]).float()[:, :, None]
    hd = torch.empty(B, S, C, device=DEV)
    capi.call("ctts_add_positions", h.to(DEV), pe.to(DEV), 2048, alpha.to(DEV), lens.to(DEV), B, S, C, 0, hd, stream())
    close(hd, ref, atol=1e-6)


def test_decode_durations_half_to_even():
    logd = torch.log(torch.tensor([1.5, 2.5, 3.5, 4.4999, 0.2, 9.0, 1.0, 1.49]) + 1.0)
    out = torch.empty_like(logd, device=DEV)
    capi.call("ctts_decode_durations", logd.to(DEV), 1.0, logd.numel(), out, stream())
    ref = torch.clamp(torch.round(torch.exp(logd) - 1) * 1.0, min=0)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("seed", range(6))
def test_length_regulator_bit_exact(seed):
    """LengthRegulator + dur_to_mel2ph against the oracle's loop restatement: bit-exact, incl. zero durations,
    fractional (d_control != 1) durations where LR truncates but mel2ph rounds, all-zero rows and max_len padding."""
    gen = g(100 + seed)
    B, S, C = 5, 37, 64
    x = torch.randn(B, S, C, generator=gen)
    src_lens = torch.tensor([37, 30, 22, 9, 1])
    if seed % 2 == 0:
        dur = torch.randint(0, 9, (B, S), generator=gen)
    else:
        dur = torch.clamp(torch.round(torch.rand(B, S, generator=gen) * 8) * (0.5 + 0.25 * seed), min=0)
    dur = dur * (torch.arange(S)[None] < src_lens[:, None])
    if seed == 2:
        dur[3] = 0
    ref_x, ref_len = O.length_regulate(x, dur, None)
    ref_m2p = O.durations_to_mel2ph(dur, torch.arange(S)[None] >= src_lens[:, None])
    out, mel_len, m2p, _ = engine.length_regulate(x.to(DEV), dur.to(DEV), src_lens.to(DEV), None, True)
    assert torch.equal(mel_len.cpu(), ref_len)
    assert torch.equal(out.cpu(), ref_x)
    assert torch.equal(m2p.cpu(), ref_m2p)
    M = int(ref_len.max()) + 11
    out2, _, _, _ = engine.length_regulate(x.to(DEV), dur.to(DEV), src_lens.to(DEV), M, False)
    ref2, _ = O.length_regulate(x, dur, M)
    assert torch.equal(out2.cpu(), ref2)


def test_length_regulator_full_size_properties():
    """BASELINE full size (B 16, S 100, 8 frames/phoneme): size-independent properties -- every output row is a copy
    of the row its mel2ph names, row counts per phoneme equal the durations, rows past mel_len are zero."""
    B, S, C = 16, 100, 256
    x = torch.randn(B, S, C, generator=g(5), device="cpu").to(DEV)
    src_lens = torch.tensor([100 - 2 * b for b in range(B)], device=DEV)
    dur = (torch.arange(S, device=DEV)[None] < src_lens[:, None]).float() * 8
    out, mel_len, m2p, _ = engine.length_regulate(x, dur, src_lens, None, True)
    assert out.shape == (B, 800, C) and mel_len.tolist() == [8 * int(s) for s in src_lens]
    idx = (m2p - 1).clamp(min=0)
    gathered = torch.gather(x, 1, idx[:, :, None].expand(-1, -1, C)) * (m2p > 0)[:, :, None]
    assert torch.equal(out, gathered)
    counts = torch.zeros(B, S + 1, dtype=torch.long, device=DEV).scatter_add(1, m2p, torch.ones_like(m2p))[:, 1:]
    assert torch.equal(counts, dur.long())


def test_cwt_to_pitch_and_buckets():
    B, T = 3, 211
    cwt = torch.randn(B, T, 11, generator=g(20))
    mean = torch.tensor([5.3, 5.0, 5.6])
    std = torch.tensor([0.4, 0.3, 0.5])
    cfg = dict(pitch_norm="log", pitch_norm_eps=1e-9, use_uv=True)
    f0n = O.cwt_to_f0_norm(cwt[:, :, :10], mean, std * 0.8, T, cfg)
    uv = cwt[:, :, -1] > 0
    f0d = O.denorm_f0(f0n, uv, cfg)
    idx_ref = O.f0_to_coarse(f0d)
    w = ((torch.arange(0, 10).float() + 1 + 2.5) ** (-2.5)).to(DEV)
    f0n_d = torch.empty(B, T, device=DEV)
    f0d_d = torch.empty(B, T, device=DEV)
    idx = torch.empty(B, T, dtype=torch.long, device=DEV)
    capi.call("ctts_cwt_to_pitch", cwt.to(DEV), 11, w, mean.to(DEV), std.to(DEV), 1, 0.8, 1e-9, None, 1, B, T, f0n_d,
              f0d_d, idx, stream())
    close(f0n_d, f0n, atol=1e-5)
    close(f0d_d, f0d, atol=1e-3, rtol=1e-5)
    flips = int((idx.cpu() != idx_ref).sum())
    assert flips == 0, "%d pitch-bucket flips" % flips
    assert len(idx_ref.unique()) > 20
    # teacher-forced variant: uv from targets, stats stride 1
    uvt = (torch.rand(B, T, generator=g(21)) < 0.3).float()
    capi.call("ctts_cwt_to_pitch", cwt[:, :, :10].contiguous().to(DEV), 10, w, mean.to(DEV), std.to(DEV), 1, 1.0, 1e-9,
              uvt.to(DEV), 1, B, T, f0n_d, f0d_d, idx, stream())
    f0n2 = O.cwt_to_f0_norm(cwt[:, :, :10], mean, std, T, cfg)
    assert int((idx.cpu() != O.f0_to_coarse(O.denorm_f0(f0n2, uvt, cfg))).sum()) == 0


def test_bucketize_and_gather():
    bins = torch.linspace(-1.43, 8.18, 255)
    v = torch.cat([torch.randn(1000, generator=g(30)) * 2 + 2, bins[::17], torch.tensor([-5.0, 20.0])])
    idx = torch.empty(v.numel(), dtype=torch.long, device=DEV)
    capi.call("ctts_bucketize", v.to(DEV), 1.0, bins.to(DEV), 255, v.numel(), idx, stream())
    assert torch.equal(idx.cpu(), torch.bucketize(v, bins))
    table = torch.randn(256, 64, generator=g(31))
    x = torch.randn(v.numel(), 64, generator=g(32))
    xd = x.to(DEV).clone()
    capi.call("ctts_gather_add", table.to(DEV), idx, v.numel(), 64, 256, xd, stream())
    assert torch.equal(xd.cpu(), x + table[idx.cpu()])


# ---------------------------------------------------------------------------------------------------------------
# tensor-core engine (tcgen05, bf16 hi/lo operand planes, 3 MMAs per k-slice)
# ---------------------------------------------------------------------------------------------------------------
def test_split_bf16_planes():
    x = torch.randn(4, 33, 64, generator=g(40)) * 3
    p = engine.split_planes(x.to(DEV))
    hi, lo = p.p[0].float().cpu(), p.p[1].float().cpu()
    assert torch.equal(hi, x.bfloat16().float())
    assert torch.equal(lo, (x - hi).bfloat16().float())
    assert (x - hi - lo).abs().max() <= x.abs().max() * 2.0 ** -16


@pytest.mark.parametrize("B,T,Cin,N,taps,act", [
    (2, 128, 256, 128, 1, "none"),      # one tile, no halo
    (2, 100, 256, 768, 1, "none"),      # T < BLOCK_M, 256-wide N tiles
    (3, 300, 256, 1024, 9, "gelu"),     # the decoder FFN conv: halo via TMA OOB fill, 36 k-blocks
    (2, 261, 1024, 256, 1, "none"),     # FFN second GEMM
    (2, 300, 80, 512, 5, "tanh"),       # PostNet first conv: Cin not a multiple of 64
    (2, 129, 512, 80, 5, "none"),       # PostNet last conv: N < BLOCK_N
    (1, 70, 256, 80, 1, "none"),        # mel_linear
    (2, 257, 512, 512, 5, "tanh"),
    (3, 96, 256, 256, 3, "relu"),       # packed row tiling, 32-row segments: tiles straddle utterance boundaries
    (3, 192, 256, 1024, 9, "gelu"),     # packed, 64-row segments, conv halo across the boundary must stay zero
    (5, 160, 512, 512, 5, "tanh"),      # packed, odd tile count (cluster padding CTA), last tile partial
    (16, 800, 256, 256, 1, "none"),     # the bench row count (100 packed tiles instead of 112)
])
def test_gemm_bf16x3_matches_fp32(B, T, Cin, N, taps, act):
    x = torch.randn(B, T, Cin, generator=g(41))
    w = torch.randn(N, Cin, taps, generator=g(42)) / math.sqrt(Cin * taps)
    bias = torch.randn(N, generator=g(43))
    res = torch.randn(B, T, N, generator=g(44))
    lens = torch.tensor([max(T - 9 * b, 1) for b in range(B)])
    sc, sh = torch.rand(N, generator=g(45)) + 0.5, torch.randn(N, generator=g(46))
    alpha = 0.61
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), padding=taps // 2).transpose(1, 2) * alpha
    ref = ref * sc.double() + sh.double()
    ref = {"none": lambda v: v, "gelu": F.gelu, "tanh": torch.tanh, "relu": F.relu}[act](ref) + res.double()
    ref = (ref * (torch.arange(T)[None, :] < lens[:, None]).double()[:, :, None]).float()
    packed = w.permute(0, 2, 1).reshape(N, -1).contiguous().to(DEV)
    xp, wp = engine.split_planes(x.to(DEV)), engine.split_planes(packed)
    y, yp = engine.gemm_tc(xp, wp, bias.to(DEV), alpha=alpha, bn=(sc.to(DEV), sh.to(DEV)), act=engine._ACTS[act],
                           residual=res.to(DEV), lens=lens.to(DEV), taps=taps, want_planes=True)
    torch.cuda.synchronize()
    # bf16x3 keeps 16 mantissa bits per operand: error ~ 2^-16 * sqrt(K) * |a||b|; far below the 1e-3 path tolerance
    close(y, ref, atol=2e-4, rtol=2e-4, msg="fp32 output")
    back = yp.value()
    close(back, y, atol=1e-4, rtol=2e-5, msg="hi/lo planes of the output")


@pytest.mark.parametrize("B,T,Cin,N,taps,act", [
    (1, 70, 256, 80, 1, "none"),        # one m-tile: the pair's second CTA has no rows; N tail by TMA zero fill
    (2, 129, 512, 80, 5, "none"),       # 4 m-tiles per-utterance tiling, conv halo
    (3, 192, 256, 1024, 9, "gelu"),     # packed 64-row segments, tiles straddle utterances
    (5, 160, 512, 512, 5, "tanh"),      # packed, odd m-tile count
    (16, 800, 256, 1024, 9, "gelu"),    # the bench's decoder FFN conv: 200 pair tiles on 74 pairs
])
def test_gemm_cta_pair_is_bit_identical(B, T, Cin, N, taps, act, monkeypatch):
    """The cta_group::2 kernel (256 x 256 tile per CTA pair) issues the same MMAs in the same order per accumulator
    element as the single-CTA persistent kernel: outputs must be bit-identical."""
    x = torch.randn(B, T, Cin, generator=g(141))
    w = torch.randn(N, taps * Cin, generator=g(142)) / math.sqrt(Cin * taps)
    bias = torch.randn(N, generator=g(143)).to(DEV)
    res = torch.randn(B, T, N, generator=g(144)).to(DEV)
    lens = torch.tensor([max(T - 9 * b, 1) for b in range(B)]).to(DEV)
    xp, wp = engine.split_planes(x.to(DEV)), engine.split_planes(w.to(DEV))
    out = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("CTTS_PAIR_GEMM", mode)
        y, yp = engine.gemm_tc(xp, wp, bias, act=engine._ACTS[act], residual=res, lens=lens, taps=taps, want_planes=True)
        torch.cuda.synchronize()
        out[mode] = (y, yp)
    assert torch.equal(out["0"][0], out["2"][0])
    for a, b in zip(out["0"][1].p, out["2"][1].p):
        assert torch.equal(a, b)
    ref = engine.conv_gemm(x.to(DEV), w.to(DEV), bias, act=engine._ACTS[act], residual=res, lens=lens, taps=taps)
    close(out["2"][0], ref, atol=2e-4, rtol=2e-4)


@pytest.mark.parametrize("B,T,Cin,N,taps", [(16, 100, 1024, 256, 1), (4, 100, 256, 256, 3), (2, 70, 256, 384, 1)])
def test_gemm_bf16x6_narrow_tiles_bit_identical(B, T, Cin, N, taps, monkeypatch):
    """128 x 64 tiles (small grids) vs 128 x 128: same MMA order per accumulator element."""
    x = torch.randn(B, T, Cin, generator=g(151))
    w = torch.randn(N, taps * Cin, generator=g(152)) / math.sqrt(Cin * taps)
    bias = torch.randn(N, generator=g(153)).to(DEV)
    lens = torch.tensor([max(T - 9 * b, 1) for b in range(B)]).to(DEV)
    xp, wp = engine.split_planes(x.to(DEV), 3), engine.split_planes(w.to(DEV), 3)
    out = {}
    for mode in ("0", "74"):           # never / the default threshold (128-wide grids of <= 74 CTAs are narrowed)
        monkeypatch.setenv("CTTS_NARROW_TILES", mode)
        y, yp = engine.gemm_tc(xp, wp, bias, act=engine._ACTS["relu"], lens=lens, taps=taps, want_planes=True)
        torch.cuda.synchronize()
        out[mode] = (y, yp)
    assert torch.equal(out["0"][0], out["74"][0])
    for a, b in zip(out["0"][1].p, out["74"][1].p):
        assert torch.equal(a, b)


def test_gemm_bf16x3_full_size_linearity():
    """BASELINE full size (B 16, T 800, FFN conv shape): size-independent property -- the op is linear in x before the
    activation, and rows beyond T / lens never leak (checked against the FP32 CUDA-core kernel on a row sample)."""
    B, T, Cin, N, taps = 16, 800, 256, 1024, 9
    gen = g(47)
    x1 = torch.randn(B, T, Cin, generator=gen).to(DEV)
    x2 = torch.randn(B, T, Cin, generator=gen).to(DEV)
    w = (torch.randn(N, taps * Cin, generator=gen) / math.sqrt(Cin * taps)).to(DEV)
    wp = engine.split_planes(w)
    y1, _ = engine.gemm_tc(engine.split_planes(x1), wp, taps=taps)
    y2, _ = engine.gemm_tc(engine.split_planes(x2), wp, taps=taps)
    y12, _ = engine.gemm_tc(engine.split_planes(x1 + x2), wp, taps=taps)
    assert (y12 - (y1 + y2)).abs().max().item() < 5e-4
    ref = engine.conv_gemm(x1, w, taps=taps)
    assert (y1 - ref).abs().max().item() < 5e-4


@pytest.mark.parametrize("n_planes", [2, 3])
@pytest.mark.parametrize("B,T,C,H", [(2, 128, 256, 2), (3, 300, 256, 2), (2, 273, 256, 2), (1, 70, 256, 4), (2, 801, 256, 2),
                                     (16, 100, 256, 2)])
def test_attention_bf16x3(B, T, C, H, n_planes):
    """Tensor-core attention (materialised scores, bf16x3 / bf16x6 GEMMs) against fp64 softmax attention."""
    qkv = torch.randn(B, T, 3 * C, generator=g(50))
    lens = torch.tensor([max(T - 31 * b, 1) for b in range(B)])
    dh = C // H
    q, k, v = qkv.double().split(C, -1)
    q = q.view(B, T, H, dh).transpose(1, 2) / math.sqrt(dh)
    k = k.view(B, T, H, dh).transpose(1, 2)
    v = v.view(B, T, H, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    pad = torch.arange(T)[None, :] >= lens[:, None]
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, C)
    ref = (ref * (~pad).double()[:, :, None]).float()
    out = engine.attention_tc(engine.split_planes(qkv.to(DEV), n_planes), lens.to(DEV), H)
    got = out.value()
    tol = 1e-4 if n_planes == 2 else 5e-6
    close(got, ref, atol=tol, rtol=tol)
    # and against the FP32 CUDA-core kernel
    close(got, engine.attention(qkv.to(DEV), lens.to(DEV), H), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("B,T,Cin,masked", [(16, 800, 256, False), (16, 800, 1024, True), (3, 100, 256, True), (2, 333, 1024, False),
                                            (5, 96, 256, False), (1, 1, 256, True)])
def test_gemm_residual_layernorm_fused(B, T, Cin, masked):
    """ctts_gemm_split_ln (projection + residual + LayerNorm in the epilogue of the 128 x 256 tile) against the two launches it
    replaces and against fp64; the residual stream is updated in place."""
    N = 256
    gen = g(81)
    x = torch.randn(B, T, Cin, generator=gen)
    w = torch.randn(N, Cin, generator=gen) / math.sqrt(Cin)
    bias = torch.randn(N, generator=gen)
    res = torch.randn(B, T, N, generator=gen)
    gamma, beta = torch.rand(N, generator=gen) + 0.5, torch.randn(N, generator=gen)
    lens = torch.tensor([max(T - 41 * b, 1) for b in range(B)])
    keep = (torch.arange(T)[None, :] < lens[:, None])[:, :, None]
    y64 = (x.double() @ w.double().t() + bias.double() + res.double()) * keep.double()
    ln64 = F.layer_norm(y64, (N,), gamma.double(), beta.double(), 1e-5)
    if masked:
        ln64 = ln64 * keep.double()
    xp, wp = engine.split_planes(x.to(DEV)), engine.split_planes(w.to(DEV))
    stream = res.to(DEV).clone()
    dl = lens.to(DEV)
    ln_y, ln_p = engine.gemm_tc_ln(xp, wp, bias.to(DEV), stream, dl, stream, gamma.to(DEV), beta.to(DEV), 1e-5,
                                   ln_lens=dl if masked else None, want_fp32=True)
    torch.cuda.synchronize()
    # the two launches it replaces
    ref_stream = res.to(DEV).clone()
    engine.gemm_tc(xp, wp, bias.to(DEV), residual=ref_stream, lens=lens.to(DEV), out=ref_stream)
    ref_y, ref_p = engine.layernorm_planes(ref_stream, gamma.to(DEV), beta.to(DEV), 1e-5, lens.to(DEV) if masked else None,
                                           want_fp32=True)
    close(stream, ref_stream, atol=1e-6, rtol=1e-6)
    close(stream, y64.float(), atol=2e-4, rtol=2e-4)
    close(ln_y, ref_y, atol=2e-5, rtol=2e-5)
    close(ln_y, ln64.float(), atol=5e-4, rtol=5e-4)
    assert (ln_p.value() - ln_y).abs().max().item() < 1e-4 * max(1.0, ln_y.abs().max().item())
    close(ln_p.value(), ref_p.value(), atol=2e-5, rtol=2e-5)


@pytest.mark.parametrize("B,T,H", [(1, 1, 2), (2, 7, 2), (3, 63, 2), (3, 64, 2), (3, 65, 2), (16, 100, 2), (4, 127, 4), (5, 128, 2)])
def test_attention_small_fused(B, T, H, monkeypatch):
    """ctts_attention_small (T <= 128, head_dim 128, 3 planes: one CTA per (batch, head)) against fp64 softmax attention and
    against the four-launch path it replaces; ragged lengths down to 1, padded query rows exactly zero."""
    C = 128 * H
    qkv = torch.randn(B, T, 3 * C, generator=g(71)) * 1.5
    lens = torch.tensor([max(T - 23 * b, 1) for b in range(B)])
    dh = C // H
    q, k, v = qkv.double().split(C, -1)
    q = q.view(B, T, H, dh).transpose(1, 2) / math.sqrt(dh)
    k = k.view(B, T, H, dh).transpose(1, 2)
    v = v.view(B, T, H, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    pad = torch.arange(T)[None, :] >= lens[:, None]
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, C)
    ref = (ref * (~pad).double()[:, :, None]).float()
    planes = engine.split_planes(qkv.to(DEV), 3)
    launches = capi.LAUNCHES
    monkeypatch.setattr(engine, "SMALL_ATTENTION", True)
    out = engine.attention_tc(planes, lens.to(DEV), H).value()
    assert capi.LAUNCHES - launches == 1, "the fused entry point was not taken"
    torch.cuda.synchronize()
    close(out, ref, atol=5e-6, rtol=5e-6)
    assert bool((out.cpu()[pad] == 0).all())
    monkeypatch.setattr(engine, "SMALL_ATTENTION", False)
    unfused = engine.attention_tc(planes, lens.to(DEV), H).value()
    close(out, unfused, atol=2e-6, rtol=2e-6)


@pytest.mark.parametrize("B,T,Cin,N,taps,act", [
    (2, 100, 256, 768, 1, "none"), (3, 100, 256, 1024, 9, "gelu"), (2, 300, 128, 256, 5, "relu"),
    (2, 261, 1024, 256, 1, "none"), (2, 100, 256, 256, 3, "relu"),
])
def test_gemm_bf16x6_is_fp32_equivalent(B, T, Cin, N, taps, act):
    """3-plane operands (24 mantissa bits, 6 MMAs per k-slice): as close to fp64 as the FP32 CUDA-core kernel is."""
    x = torch.randn(B, T, Cin, generator=g(61))
    w = torch.randn(N, Cin, taps, generator=g(62)) / math.sqrt(Cin * taps)
    bias = torch.randn(N, generator=g(63))
    lens = torch.tensor([max(T - 9 * b, 1) for b in range(B)])
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), padding=taps // 2).transpose(1, 2)
    ref = {"none": lambda v: v, "gelu": F.gelu, "relu": F.relu}[act](ref)
    ref = (ref * (torch.arange(T)[None, :] < lens[:, None]).double()[:, :, None]).float()
    packed = w.permute(0, 2, 1).reshape(N, -1).contiguous().to(DEV)
    y, yp = engine.gemm_tc(engine.split_planes(x.to(DEV), 3), engine.split_planes(packed, 3), bias.to(DEV),
                           act=engine._ACTS[act], lens=lens.to(DEV), taps=taps, want_planes=True)
    y32 = engine.conv_gemm(x.to(DEV), packed, bias.to(DEV), act=engine._ACTS[act], lens=lens.to(DEV), taps=taps)
    err6 = (y.cpu() - ref).abs().max().item()
    err32 = (y32.cpu() - ref).abs().max().item()
    print('bf16x6 err %.3g  fp32 err %.3g' % (err6, err32))
    assert err6 <= max(3 * err32, 8e-6), (err6, err32)
    assert (yp.value() - y).abs().max().item() < 1e-6
    assert yp.n == 3


@pytest.mark.parametrize("B,T,H", [(2, 128, 2), (3, 300, 2), (2, 273, 2), (2, 64, 2), (2, 801, 2), (16, 800, 2), (1, 1000, 4)])
def test_flash_attention_bf16x3(B, T, H):
    """Fused tensor-core attention (S in TMEM, P planes in shared memory) against fp64 and against the materialised
    tensor-core path."""
    C = 128 * H
    qkv = torch.randn(B, T, 3 * C, generator=g(70)) * 1.5
    lens = torch.tensor([max(T - 37 * b, 1) for b in range(B)])
    dh = C // H
    q, k, v = qkv.double().split(C, -1)
    q = q.view(B, T, H, dh).transpose(1, 2) / math.sqrt(dh)
    k = k.view(B, T, H, dh).transpose(1, 2)
    v = v.view(B, T, H, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    pad = torch.arange(T)[None, :] >= lens[:, None]
    s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, C)
    ref = (ref * (~pad).double()[:, :, None]).float()
    planes = engine.split_planes(qkv.to(DEV), 2)
    out = engine.attention_flash(planes, lens.to(DEV), H)
    torch.cuda.synchronize()
    close(out.value(), ref, atol=1e-4, rtol=1e-4)
