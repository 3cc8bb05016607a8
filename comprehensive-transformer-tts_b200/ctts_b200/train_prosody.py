"""liu2021 implicit prosody modelling in TRAINING mode (modules.py:1002-1023): the reference encoders
(ReferenceEncoder modules.py:332-397 with CoordConv2d coordconv.py:140-159 + BatchNorm2d on batch statistics,
UtteranceLevelProsodyEncoder :537-569 with the style-token attention :453-534, PhonemeLevelProsodyEncoder :400-450) and
the parallel predictors (:572-648) on the tape of train_engine.py.

The 3x3 / stride (1, 2) convolutions of the reference encoder run as im2col + the dense GEMM engine on channels-last
activations; the small attention products run on the strided FP32 GEMM."""
import math

import torch

from . import capi
from .capi import ACT_NONE, ACT_RELU, ACT_TANH
from . import train_engine as TE
from .train_engine import (Var, accumulate_into, add_row, batch_norm_act, dropout, grad_buffer, layer_norm, linear,
                           _generic, _reshape, _st)


# ---------------------------------------------------------------------------------------------------------------------
# small tape ops
# ---------------------------------------------------------------------------------------------------------------------
def mm_nt(ctx, a, b, alpha=1.0):
    """y[z] = alpha * a[z] b[z]^T;  a [Z, M, K], b [Z, N, K] (b may have Z = 1: shared)."""
    Z, M, K = a.v.shape
    Zb, N, _ = b.v.shape
    bz = N * K if Zb == Z else 0
    y = Var(torch.empty(Z, M, N, device=a.v.device, dtype=torch.float32))
    _generic(a.v, b.v, y.v, Z, 1, M, N, K, (M * K, 0, K, 1, 0), (bz, 0, K, 1, 0), (M * N, 0, N, 1), alpha=alpha)

    def bwd():
        if y.g is None:
            return
        if a.needs_grad:      # da[m, k] = alpha * sum_n dy[m, n] b[n, k]
            da, acc = grad_buffer(a)
            _generic(y.g, b.v, da, Z, 1, M, K, N, (M * N, 0, N, 1, 0), (bz, 0, 1, K, 0), (M * K, 0, K, 1), alpha=alpha,
                     accumulate=acc)
        if b.needs_grad:      # db[n, k] = alpha * sum_m dy[m, n] a[m, k]   (summed over z when b is shared)
            db, acc = grad_buffer(b)
            if Zb == Z:
                _generic(y.g, a.v, db, Z, 1, N, K, M, (M * N, 0, 1, N, 0), (M * K, 0, 1, K, 0), (N * K, 0, K, 1), alpha=alpha,
                         accumulate=acc)
            else:
                _generic(y.g, a.v, db, 1, 1, N, K, Z * M, (0, 0, 1, N, 0), (0, 0, 1, K, 0), (0, 0, K, 1), alpha=alpha,
                         accumulate=acc)
        y.g = None

    ctx.record(bwd)
    return y


def mm_nn(ctx, a, b):
    """y[z] = a[z] b[z];  a [Z, M, K], b [Z, K, N] (b may have Z = 1: shared)."""
    Z, M, K = a.v.shape
    Zb, _, N = b.v.shape
    bz = K * N if Zb == Z else 0
    y = Var(torch.empty(Z, M, N, device=a.v.device, dtype=torch.float32))
    _generic(a.v, b.v, y.v, Z, 1, M, N, K, (M * K, 0, K, 1, 0), (bz, 0, 1, N, 0), (M * N, 0, N, 1))

    def bwd():
        if y.g is None:
            return
        if a.needs_grad:      # da[m, k] = sum_n dy[m, n] b[k, n]
            da, acc = grad_buffer(a)
            _generic(y.g, b.v, da, Z, 1, M, K, N, (M * N, 0, N, 1, 0), (bz, 0, N, 1, 0), (M * K, 0, K, 1), accumulate=acc)
        if b.needs_grad:      # db[k, n] = sum_m a[m, k] dy[m, n]
            db, acc = grad_buffer(b)
            if Zb == Z:
                _generic(a.v, y.g, db, Z, 1, K, N, M, (M * K, 0, 1, K, 0), (M * N, 0, 1, N, 0), (K * N, 0, N, 1), accumulate=acc)
            else:
                _generic(a.v, y.g, db, 1, 1, K, N, Z * M, (0, 0, 1, K, 0), (0, 0, 1, N, 0), (0, 0, N, 1), accumulate=acc)
        y.g = None

    ctx.record(bwd)
    return y


def softmax_rows(ctx, s, key_lens=None):
    """softmax over the last dim of s [Z, T, Tk]; keys >= key_lens[z] excluded."""
    Z, T, Tk = s.v.shape
    y = Var(torch.empty_like(s.v))
    capi.call("ctts_masked_softmax", s.v, key_lens, 1, Z, T, Tk, Tk, 0, y.v, _st())

    def bwd():
        if y.g is None:
            return
        ds = torch.empty_like(s.v)
        capi.call("ctts_softmax_bwd", y.v, y.g, Z, T, Tk, Tk, 1.0, ds, _st())
        accumulate_into(s, ds)
        y.g = None

    ctx.record(bwd)
    return y


def mask_rows(ctx, x, lens):
    """x[b, t, :] zeroed for t >= lens[b] (masked_fill(mask.unsqueeze(-1), 0))."""
    B, T, C = x.v.shape
    y = Var(x.v.clone())
    capi.call("ctts_mask_rows", y.v, lens, B, T, C, _st())

    def bwd():
        if y.g is None:
            return
        capi.call("ctts_mask_rows", y.g, lens, B, T, C, _st())
        accumulate_into(x, y.g)
        y.g = None

    ctx.record(bwd)
    return y


def slice_cols(ctx, x, c0, C):
    """x[..., c0:c0+C] as a contiguous tensor."""
    lead = x.v.shape[:-1]
    ld = x.v.shape[-1]
    rows = x.v.numel() // ld
    y = Var(torch.empty(*lead, C, device=x.v.device, dtype=torch.float32))
    capi.call("ctts_copy_rows", x.v.view(-1)[c0:], ld, rows, C, y.v, C, 0, _st())

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        if x.g is None:
            x.g = torch.zeros_like(x.v)
        capi.call("ctts_copy_rows", y.g, C, rows, C, x.g.view(-1)[c0:], ld, 1, _st())
        y.g = None

    ctx.record(bwd)
    return y


def tanh_param(ctx, name):
    """tanh of a parameter (the style-token table, modules.py:484)."""
    w = ctx.P[name]
    y = Var(torch.empty_like(w))
    capi.call("ctts_act_fwd", w, w.numel(), ACT_TANH, y.v, 0, None, _st())

    def bwd():
        G = ctx.G.get(name)
        if y.g is None or G is None:
            return
        capi.call("ctts_act_bwd", y.g, y.v, ACT_TANH, 1.0, None, 1, 1, 1, w.numel(), y.g, None, _st())
        capi.call("ctts_axpy", y.g, 1.0, w.numel(), 1, G, _st())
        y.g = None

    ctx.record(bwd)
    return y


# ---------------------------------------------------------------------------------------------------------------------
# GRU layers
# ---------------------------------------------------------------------------------------------------------------------
def _gru_param_grads(ctx, pre, suffix, dgh, out, out_ld, out_off, B, T, H, reverse):
    """dW_hh = sum_t dgh_t (x) h_{t-1}, db_hh = sum dgh  (h_{t-1} = `out` one step earlier in processing order)."""
    Gw, Gb = ctx.G.get(pre + "weight_hh_l0" + suffix), ctx.G.get(pre + "bias_hh_l0" + suffix)
    if Gw is not None:
        src = out.view(-1)[out_off:]
        _generic(dgh, src, Gw, 1, 1, 3 * H, H, B * T, (0, 0, 1, 3 * H, T * 3 * H), (0, 0, 1, out_ld, T * out_ld), (0, 0, H, 1),
                 Kin=T, shift0=1 if reverse else -1, accumulate=1)
    if Gb is not None:
        capi.call("ctts_act_bwd", dgh, None, ACT_NONE, 1.0, None, 1, T, B * T, 3 * H, None, Gb, _st())


def gru_uni(ctx, x, pre):
    """Single-layer unidirectional nn.GRU over all T steps: (memory [B,T,H], last hidden state [B,H])."""
    P = ctx.P
    B, T, _ = x.v.shape
    H = P[pre + "weight_hh_l0"].shape[1]
    gi = linear(ctx, x, pre + "weight_ih_l0", pre + "bias_ih_l0")
    st = _st()
    dev = x.v.device
    out2 = torch.empty(B, T, 2 * H, device=dev, dtype=torch.float32)
    hf2 = torch.empty(B, 2 * H, device=dev, dtype=torch.float32)
    whh, bhh = P[pre + "weight_hh_l0"], P[pre + "bias_hh_l0"]
    capi.call("ctts_gru_bidir", gi.v, gi.v, whh, bhh, whh, bhh, B, T, H, out2, hf2, st)    # the reverse half is unused
    mem = Var(torch.empty(B, T, H, device=dev, dtype=torch.float32))
    last = Var(torch.empty(B, H, device=dev, dtype=torch.float32))
    capi.call("ctts_copy_rows", out2, 2 * H, B * T, H, mem.v, H, 0, st)
    capi.call("ctts_copy_rows", hf2, 2 * H, B, H, last.v, H, 0, st)

    def bwd():
        if mem.g is None and last.g is None:
            return
        dgi = torch.empty_like(gi.v)
        dgh = torch.empty_like(gi.v)
        capi.call("ctts_gru_bwd", gi.v, whh, bhh, mem.v, H, 0, mem.g, last.g, H, B, T, H, 0, dgi, dgh, _st())
        _gru_param_grads(ctx, pre, "", dgh, mem.v, H, 0, B, T, H, 0)
        accumulate_into(gi, dgi)
        mem.g = last.g = None

    ctx.record(bwd)
    return mem, last


def gru_bidir(ctx, x, pre):
    """Bidirectional nn.GRU: (outputs [B,T,2H] = fwd | bwd, final states [B,2H])."""
    P = ctx.P
    B, T, _ = x.v.shape
    H = P[pre + "weight_hh_l0"].shape[1]
    gi_f = linear(ctx, x, pre + "weight_ih_l0", pre + "bias_ih_l0")
    gi_b = linear(ctx, x, pre + "weight_ih_l0_reverse", pre + "bias_ih_l0_reverse")
    dev = x.v.device
    out = Var(torch.empty(B, T, 2 * H, device=dev, dtype=torch.float32))
    hfin = Var(torch.empty(B, 2 * H, device=dev, dtype=torch.float32))
    capi.call("ctts_gru_bidir", gi_f.v, gi_b.v, P[pre + "weight_hh_l0"], P[pre + "bias_hh_l0"],
              P[pre + "weight_hh_l0_reverse"], P[pre + "bias_hh_l0_reverse"], B, T, H, out.v, hfin.v, _st())

    def bwd():
        if out.g is None and hfin.g is None:
            return
        for rev, gi, suffix in ((0, gi_f, ""), (1, gi_b, "_reverse")):
            dgi = torch.empty_like(gi.v)
            dgh = torch.empty_like(gi.v)
            dhf = hfin.g.view(-1)[rev * H:] if hfin.g is not None else None
            capi.call("ctts_gru_bwd", gi.v, P[pre + "weight_hh_l0" + suffix], P[pre + "bias_hh_l0" + suffix], out.v, 2 * H,
                      rev * H, out.g, dhf, 2 * H, B, T, H, rev, dgi, dgh, _st())
            _gru_param_grads(ctx, pre, suffix, dgh, out.v, 2 * H, rev * H, B, T, H, rev)
            accumulate_into(gi, dgi)
        out.g = hfin.g = None

    ctx.record(bwd)
    return out, hfin


# ---------------------------------------------------------------------------------------------------------------------
# reference encoder
# ---------------------------------------------------------------------------------------------------------------------
def _conv3x3_s12(ctx, x4, wname, bname):
    """Conv2d(3x3, stride (1, 2), pad (1, 1)) on channels-last x4 [N, H, W, C] -> Var [1, N*H*Wo, Cout]."""
    N, H, W, C = x4.v.shape
    Wo = (W + 2 - 3) // 2 + 1
    dev = x4.v.device
    w = ctx.P[wname]
    Cout = w.shape[0]
    col = Var(torch.empty(1, N * H * Wo, 9 * C, device=dev, dtype=torch.float32), x4.needs_grad)
    capi.call("ctts_im2col_3x3_s12", x4.v, N, H, W, C, col.v, _st())

    def bwd_col():
        if col.g is None or not x4.needs_grad:
            return
        dx = torch.empty_like(x4.v)
        capi.call("ctts_col2im_3x3_s12", col.g, N, H, W, C, dx, _st())
        accumulate_into(x4, dx)
        col.g = None

    ctx.record(bwd_col)
    # the GEMM weight: [Cout, (kh*3 + kw)*Cin + c] = the tap-major packing of the [Cout, Cin, 9] view
    vname = wname + "#im2col"
    ctx.src[vname] = [w]

    def build():
        out = torch.empty(Cout, 9 * C, device=dev, dtype=torch.float32)
        capi.call("ctts_pack_conv_weight", w.detach().reshape(Cout, C, 9), Cout, C, 9, out, _st())
        return out

    ctx.P[vname] = ctx.tw.get(("im2col_w", wname), [w], build)
    holder = {}

    def distribute():
        t = holder.get("g")
        G = ctx.G.get(wname)
        if t is not None and G is not None:
            capi.call("ctts_unpack_conv_wgrad", t, Cout, C, 9, 1, G, _st())
        ctx.G.pop(vname, None)

    ctx.record(distribute)
    z = linear(ctx, col, vname, bname)

    def alloc():
        holder["g"] = ctx.G[vname] = torch.zeros(Cout, 9 * C, device=dev, dtype=torch.float32)

    ctx.record(alloc)
    return z, Wo


def reference_encoder(ctx, pre, mel, mel_lens):
    """ReferenceEncoder.forward (modules.py:370-392): CoordConv2d + 5 Conv2d, each + BatchNorm2d (batch statistics) + ReLU,
    then a GRU over time.  mel [N, M, n_mel] (no gradient).  Returns (memory [N, M, g], last hidden state [N, g])."""
    c = ctx.cfg["prosody_modeling"]["liu2021"]
    assert tuple(c["ref_enc_size"]) == (3, 3) and tuple(c["ref_enc_strides"]) == (1, 2) and tuple(c["ref_enc_pad"]) == (1, 1), \
        "reference encoder: only the shipped 3x3 / stride (1, 2) / pad (1, 1) geometry is built"
    N, M, n_mel = mel.shape
    dev = mel.device
    x0 = torch.empty(N, M, n_mel, 4, device=dev, dtype=torch.float32)
    capi.call("ctts_add_coords", mel, N, M, n_mel, x0, _st())
    x4 = Var(x0, False)
    W = n_mel
    for i in range(len(c["ref_enc_filters"])):
        wname = pre + ("convs.0.conv.weight" if i == 0 else "convs.%d.weight" % i)
        bname = pre + ("convs.0.conv.bias" if i == 0 else "convs.%d.bias" % i)
        z, W = _conv3x3_s12(ctx, x4, wname, bname)
        y = batch_norm_act(ctx, z, pre + "bns.%d." % i, ACT_RELU)
        x4 = _reshape(ctx, y, (N, M, W, z.v.shape[-1]))
    # [N, M, W, C] -> the reference's [N, M, C*W] (channel-major: out.transpose(1, 2).view(N, T, -1) of an NCHW tensor)
    Cc = x4.v.shape[-1]
    perm = Var(torch.empty(N, M, Cc * W, device=dev, dtype=torch.float32))
    capi.call("ctts_permute_last2", x4.v, N * M, W, Cc, perm.v, _st())

    def bwd_perm():
        if perm.g is None:
            return
        dx = torch.empty_like(x4.v)
        capi.call("ctts_permute_last2", perm.g, N * M, Cc, W, dx, _st())
        accumulate_into(x4, dx)
        perm.g = None

    ctx.record(bwd_perm)
    seq = mask_rows(ctx, perm, mel_lens)
    return gru_uni(ctx, seq, pre + "gru.")


def utterance_prosody_encoder(ctx, mel, mel_lens):
    """UtteranceLevelProsodyEncoder.forward (modules.py:555-569) with the single-head style-token attention (:471-533)."""
    pre = "variance_adaptor.utterance_prosody_encoder."
    E = ctx.cfg["transformer"]["encoder_hidden"]
    N = mel.shape[0]
    _, last = reference_encoder(ctx, pre + "encoder.", mel, mel_lens)
    query = linear(ctx, _reshape(ctx, last, (1, N, -1)), pre + "encoder_prj.weight", pre + "encoder_prj.bias")    # [1, N, E/2]
    tokens = _reshape(ctx, tanh_param(ctx, pre + "stl.embed"), (1, -1, E))                                          # [1, n_tok, E]
    values = linear(ctx, tokens, pre + "stl.attention.W_value.weight")
    querys = linear(ctx, query, pre + "stl.attention.W_query.weight")                                               # [1, N, E]
    keys = linear(ctx, tokens, pre + "stl.attention.W_key.weight")
    scores = softmax_rows(ctx, mm_nt(ctx, querys, keys, alpha=1.0 / math.sqrt(E)))                                  # [1, N, n_tok]
    style = mm_nn(ctx, scores, values)                                                                               # [1, N, E]
    out = linear(ctx, style, pre + "encoder_bottleneck.weight", pre + "encoder_bottleneck.bias")
    return _reshape(ctx, out, (N, 1, -1))


def phoneme_prosody_encoder(ctx, x, src_lens, mel, mel_lens):
    """PhonemeLevelProsodyEncoder.forward (modules.py:421-450): text queries attend over the reference-encoder memory."""
    pre = "variance_adaptor.phoneme_prosody_encoder."
    E = ctx.cfg["transformer"]["encoder_hidden"]
    memory, _ = reference_encoder(ctx, pre + "encoder.", mel, mel_lens)
    emb = linear(ctx, memory, pre + "encoder_prj.weight", pre + "encoder_prj.bias")          # [B, M, 2E]
    k = slice_cols(ctx, emb, 0, E)
    v = slice_cols(ctx, emb, E, E)
    q = linear(ctx, x, pre + "linears.0.linear.weight")
    k = linear(ctx, k, pre + "linears.1.linear.weight")
    attn = softmax_rows(ctx, mm_nt(ctx, q, k, alpha=1.0 / math.sqrt(E)), key_lens=mel_lens)  # [B, S, M]
    attn = mask_rows(ctx, attn, src_lens)
    out = linear(ctx, mm_nn(ctx, attn, v), pre + "encoder_bottleneck.weight", pre + "encoder_bottleneck.bias", lens=src_lens)
    return out, attn


# ---------------------------------------------------------------------------------------------------------------------
def prosody_predictor(ctx, pre, x, phoneme_level):
    """ParallelProsodyPredictor.forward (modules.py:630-648): (conv k3 + ReLU + LayerNorm + dropout) x 2, bi-GRU, Linear."""
    c = ctx.cfg["prosody_modeling"]["liu2021"]
    k = c["predictor_kernel_size"]
    if k != 3:
        raise NotImplementedError("conv1d_2 uses padding=1: only predictor_kernel_size 3 is 'same' (modules.py:610)")
    B = x.v.shape[0]
    h = x
    for i in (1, 2):
        z = linear(ctx, h, pre + "conv_layer.conv1d_%d.conv.weight" % i, pre + "conv_layer.conv1d_%d.conv.bias" % i,
                   act=ACT_RELU, taps=k)
        h = layer_norm(ctx, z, pre + "conv_layer.layer_norm_%d.weight" % i, pre + "conv_layer.layer_norm_%d.bias" % i, 1e-5)
        h = dropout(ctx, h, c["predictor_dropout"])
    out, hfin = gru_bidir(ctx, h, pre + "gru.")
    if phoneme_level:
        return linear(ctx, out, pre + "predictor_bottleneck.weight", pre + "predictor_bottleneck.bias")
    y = linear(ctx, _reshape(ctx, hfin, (1, B, -1)), pre + "predictor_bottleneck.weight", pre + "predictor_bottleneck.bias")
    return _reshape(ctx, y, (B, 1, -1))


def liu2021(ctx, x, src_lens, mel, mel_lens):
    """The liu2021 branch of VarianceAdaptor.forward in training mode (modules.py:1002-1023).  Returns (x, prosody_info)."""
    pre = "variance_adaptor."
    assert mel is not None and mel_lens is not None, "liu2021 training needs the target mels"
    B = x.v.shape[0]
    mel = TE._f32(mel)
    u_emb = utterance_prosody_encoder(ctx, mel, mel_lens)                               # [B, 1, E]
    p_emb, p_attn = phoneme_prosody_encoder(ctx, x, src_lens, mel, mel_lens)            # [B, S, 4], [B, S, M]
    u_vec = prosody_predictor(ctx, pre + "utterance_prosody_predictor.", x, False)      # [B, 1, E]
    u_add = linear(ctx, _reshape(ctx, u_emb, (1, B, -1)), pre + "utterance_prosody_prj.weight",
                   pre + "utterance_prosody_prj.bias")
    x = add_row(ctx, x, _reshape(ctx, u_add, (B, -1)))
    p_vec = prosody_predictor(ctx, pre + "phoneme_prosody_predictor.", x, True)         # [B, S, 4]
    x = linear(ctx, p_emb, pre + "phoneme_prosody_prj.weight", pre + "phoneme_prosody_prj.bias", residual=x)
    return x, (u_emb, p_emb, u_vec, p_vec, p_attn)
