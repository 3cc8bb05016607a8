"""Encoder / decoder stacks of the `transformer`, `fastformer` and `conformer` block types over the C ABI.

Same conventions as engine.py (which holds the fs2 stack and the shared VarianceAdaptor / mel head): the encoder is
FP32 on the CUDA cores (it feeds the quantisers), the decoder runs its dense contractions on tcgen05 (bf16x3) when
`module.decoder_math == "bf16x3"`.  Reference: model/transformers/{transformer,fastformer,conformer}.py.
"""
import math

import numpy as np
import torch

from . import capi
from .capi import ACT_GELU, ACT_NONE, ACT_RELU
from .engine import (Planes, _stream, attention, attention_tc, conv_gemm, gemm_batched_planes, gemm_tc, layernorm,
                     layernorm_planes, split_planes)


def _interleaved_table(rows, d, device):
    """get_sinusoid_encoding_table (blocks.py:26-46): float64 on the host, then cast -- regenerated when an eval-mode
    sequence is longer than max_seq_len (transformer.py:65-70,128-135)."""
    pos = np.arange(rows, dtype=np.float64)[:, None]
    tab = pos / np.power(10000, 2 * (np.arange(d)[None, :] // 2) / d)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.FloatTensor(tab).to(device).contiguous()


def abs_table(prep, P, key, T, d, max_seq_len, device):
    if T <= max_seq_len:
        return P[key][0]
    ck = ("interleaved", d, str(device))
    tab = prep.pe.get(ck)
    if tab is None or tab.shape[0] < T:
        if tab is not None:
            prep.pe_retired.append(tab)    # captured CUDA graphs hold its address (rows are a prefix of the new table)
            T = max(T, 2 * tab.shape[0])
        tab = _interleaved_table(T, d, device)
        prep.pe[ck] = tab
    return tab


def embed_abs(prep, P, cfg, d, tokens):
    """src_word_emb(src_seq) + position_enc[:, :S]  (transformer.py:62-74; same in fastformer / conformer)."""
    B, S = tokens.shape
    table = P["encoder.src_word_emb.weight"]
    pe = abs_table(prep, P, "encoder.position_enc", S, d, cfg["max_seq_len"], tokens.device)
    x = torch.empty(B, S, d, device=tokens.device, dtype=torch.float32)
    word = torch.empty_like(x)
    capi.call("ctts_embed_tokens", tokens, table, pe, pe.shape[0], 1.0, B, S, d, table.shape[0], x, word, None, 1,
              _stream())
    return x, word


def add_abs_positions(prep, P, cfg, key, x):
    B, T, C = x.shape
    pe = abs_table(prep, P, key, T, C, cfg["max_seq_len"], x.device)
    y = torch.empty_like(x)
    capi.call("ctts_add_positions", x, pe, pe.shape[0], None, None, B, T, C, 1, y, _stream())
    return y


# ---------------------------------------------------------------------------------------------
# transformer (post-LN), model/transformers/transformer.py:157-288
# ---------------------------------------------------------------------------------------------
def prepare_transformer(prep, P, tc_decoder):
    """Concatenate w_qs | w_ks | w_vs into one [3d, d] projection (the layout ctts_attention consumes)."""
    for name in list(P):
        if name.endswith("slf_attn.w_qs.linear.weight"):
            base = name[: -len("w_qs.linear.weight")]
            cat = torch.cat([P[base + "w_qs.linear.weight"], P[base + "w_ks.linear.weight"],
                             P[base + "w_vs.linear.weight"]], 0).float().contiguous()
            prep.w[base + "qkv"] = cat
            _cat_planes(prep, base + "qkv", cat, name, tc_decoder)


def _cat_planes(prep, key, cat, name, tc_decoder):
    """Operand planes of a concatenated projection: 2 planes on the decoder (bf16x3), 3 on the encoder (bf16x6)."""
    if tc_decoder and name.startswith("decoder."):
        prep.w[key + "#planes"] = split_planes(cat)
    if prep.module.encoder_math == "bf16x6" and name.startswith("encoder."):
        prep.w[key + "#planes3"] = split_planes(cat, 3)


def _n_planes(prep, pre, tc_decoder):
    """0 = FP32 CUDA cores, 2 = bf16x3 (decoder), 3 = bf16x6 (encoder: FP32-equivalent, upstream of the quantisers)."""
    if pre.startswith("decoder."):
        return 2 if tc_decoder else 0
    return 3 if prep.module.encoder_math == "bf16x6" else 0


def _stack_transformer(prep, P, pre, x, lens, n_layers, n_head, kernel, n):
    W = prep.w
    tag = "#planes" if n == 2 else "#planes3"
    xp = split_planes(x, n) if n else None
    for i in range(n_layers):
        a = "%slayer_stack.%d.slf_attn." % (pre, i)
        f = "%slayer_stack.%d.pos_ffn." % (pre, i)
        if n:
            _, qkvp = gemm_tc(xp, W[a + "qkv" + tag], want_fp32=False, want_planes=True)
            ap = attention_tc(qkvp, lens, n_head)
            o, _ = gemm_tc(ap, W[a + "fc.linear.weight" + tag], residual=x)
            x, xp = layernorm_planes(o, P[a + "layer_norm.weight"], P[a + "layer_norm.bias"], 1e-5, lens, want_fp32=True, n=n)
            _, hp = gemm_tc(xp, W[f + "w_1.weight" + tag], P[f + "w_1.bias"], act=ACT_RELU, taps=kernel[0],
                            want_fp32=False, want_planes=True)
            o, _ = gemm_tc(hp, W[f + "w_2.weight" + tag], P[f + "w_2.bias"], residual=x, taps=kernel[1])
            x, xp = layernorm_planes(o, P[f + "layer_norm.weight"], P[f + "layer_norm.bias"], 1e-5, lens, want_fp32=True, n=n)
        else:
            qkv = conv_gemm(x, W[a + "qkv"])
            att = attention(qkv, lens, n_head)
            o = conv_gemm(att, P[a + "fc.linear.weight"], residual=x)
            x = layernorm(o, P[a + "layer_norm.weight"], P[a + "layer_norm.bias"], 1e-5, lens)
            h = conv_gemm(x, W[f + "w_1.weight"], P[f + "w_1.bias"], act=ACT_RELU, taps=kernel[0])
            o = conv_gemm(h, W[f + "w_2.weight"], P[f + "w_2.bias"], residual=x, taps=kernel[1])
            x = layernorm(o, P[f + "layer_norm.weight"], P[f + "layer_norm.bias"], 1e-5, lens)
    return x, xp


def encoder_transformer(prep, P, cfg, tokens, src_lens):
    c = cfg["transformer"]
    x, word = embed_abs(prep, P, cfg, c["encoder_hidden"], tokens)
    x, _ = _stack_transformer(prep, P, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"],
                              c["conv_kernel_size"], _n_planes(prep, "encoder.", False))
    return x, word


def decoder_transformer(prep, P, cfg, x, mel_lens, math_mode):
    c = cfg["transformer"]
    x = add_abs_positions(prep, P, cfg, "decoder.position_enc", x)
    x, xp = _stack_transformer(prep, P, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"],
                               c["conv_kernel_size"], _n_planes(prep, "decoder.", math_mode == "bf16x3"))
    return x, xp


# ---------------------------------------------------------------------------------------------
# fastformer (additive attention), model/transformers/fastformer.py:140-376
# ---------------------------------------------------------------------------------------------
def prepare_fastformer(prep, P, tc_decoder):
    pass  # every weight is consumed as stored (tap-major conv packing / planes are generic)


def _binary(a, b, op, rowwise=False, lens=None):
    B, T, C = a.shape
    y = torch.empty_like(a)
    capi.call("ctts_binary", a, b, op, 1 if rowwise else 0, lens, B, T, C, y, _stream())
    return y


def _pool(logits, values, lens, heads, hs):
    B, T, _ = logits.shape
    pooled = torch.empty(B, heads * hs, device=logits.device, dtype=torch.float32)
    capi.call("ctts_fastformer_pool", logits, values, lens, B, T, heads, hs, pooled, _stream())
    return pooled


def _stack_fastformer(prep, P, pre, x, lens, n_layers, heads, kernel, n):
    """FFTBlock.forward (fastformer.py:163-171) with FastAttention (:296-345); `heads` = the ctor's dim_head (128)."""
    W = prep.w
    C = x.shape[-1]
    hs = C // heads
    tied = "%slayer_stack.layers.0.0.fn." % pre   # to_{q,k}_attn_logits are shared by all layers (:157-161)
    tag = "#planes" if n == 2 else "#planes3"
    for i in range(n_layers):
        a = "%slayer_stack.layers.%d.0." % (pre, i)
        f = "%slayer_stack.layers.%d.1." % (pre, i)
        if n:
            _, hp = layernorm_planes(x, P[a + "norm.weight"], P[a + "norm.bias"], 1e-5, n=n)
            q, qp = gemm_tc(hp, W[a + "fn.query.weight" + tag], P[a + "fn.query.bias"], want_planes=True)
            k, _ = gemm_tc(hp, W[a + "fn.key.weight" + tag], P[a + "fn.key.bias"])
            ql, _ = gemm_tc(qp, W[tied + "to_q_attn_logits.weight" + tag], P[tied + "to_q_attn_logits.bias"])
            pooled_q = _pool(ql, q, lens, heads, hs)
            qk = _binary(k, pooled_q, 1, rowwise=True)
            kl, _ = gemm_tc(split_planes(qk, n), W[tied + "to_k_attn_logits.weight" + tag], P[tied + "to_k_attn_logits.bias"])
            pooled_k = _pool(kl, qk, lens, heads, hs)
            wv = _binary(q, pooled_k, 1, rowwise=True)
            r = _binary(q, x, 0)
            gemm_tc(split_planes(wv, n), W[a + "fn.transform.weight" + tag], P[a + "fn.transform.bias"], residual=r,
                    lens=lens, out=x)
            _, hp = layernorm_planes(x, P[f + "norm.weight"], P[f + "norm.bias"], 1e-5, n=n)
            _, gp = gemm_tc(hp, W[f + "fn.w_1.weight" + tag], P[f + "fn.w_1.bias"], act=ACT_GELU, taps=kernel[0],
                            want_fp32=False, want_planes=True)
            gemm_tc(gp, W[f + "fn.w_2.weight" + tag], P[f + "fn.w_2.bias"], residual=x, lens=lens, taps=kernel[1], out=x)
        else:
            h = layernorm(x, P[a + "norm.weight"], P[a + "norm.bias"], 1e-5)
            q = conv_gemm(h, P[a + "fn.query.weight"], P[a + "fn.query.bias"])
            k = conv_gemm(h, P[a + "fn.key.weight"], P[a + "fn.key.bias"])
            ql = conv_gemm(q, P[tied + "to_q_attn_logits.weight"], P[tied + "to_q_attn_logits.bias"])
            pooled_q = _pool(ql, q, lens, heads, hs)
            qk = _binary(k, pooled_q, 1, rowwise=True)
            kl = conv_gemm(qk, P[tied + "to_k_attn_logits.weight"], P[tied + "to_k_attn_logits.bias"])
            pooled_k = _pool(kl, qk, lens, heads, hs)
            wv = _binary(q, pooled_k, 1, rowwise=True)
            r = _binary(q, x, 0)
            conv_gemm(wv, P[a + "fn.transform.weight"], P[a + "fn.transform.bias"], residual=r, lens=lens, out=x)
            h = layernorm(x, P[f + "norm.weight"], P[f + "norm.bias"], 1e-5)
            g = conv_gemm(h, W[f + "fn.w_1.weight"], P[f + "fn.w_1.bias"], act=ACT_GELU, taps=kernel[0])
            conv_gemm(g, W[f + "fn.w_2.weight"], P[f + "fn.w_2.bias"], residual=x, lens=lens, taps=kernel[1], out=x)
    return x


def encoder_fastformer(prep, P, cfg, tokens, src_lens):
    c = cfg["transformer"]  # sic (fastformer.py:24-34)
    x, word = embed_abs(prep, P, cfg, c["encoder_hidden"], tokens)
    heads = c["encoder_hidden"] // c["encoder_head"]
    return _stack_fastformer(prep, P, "encoder.", x, src_lens, c["encoder_layer"], heads, c["conv_kernel_size"],
                             _n_planes(prep, "encoder.", False)), word


def decoder_fastformer(prep, P, cfg, x, mel_lens, math_mode):
    c = cfg["transformer"]
    x = add_abs_positions(prep, P, cfg, "decoder.position_enc", x)
    heads = c["decoder_hidden"] // c["decoder_head"]
    return _stack_fastformer(prep, P, "decoder.", x, mel_lens, c["decoder_layer"], heads, c["conv_kernel_size"],
                             _n_planes(prep, "decoder.", math_mode == "bf16x3")), None


# ---------------------------------------------------------------------------------------------
# conformer, model/transformers/conformer.py:162-560
# ---------------------------------------------------------------------------------------------
def prepare_conformer(prep, P, tc_decoder):
    for name in list(P):
        if name.endswith("sequential.2.module.sequential.5.weight"):   # eval-mode BatchNorm1d fold (conformer.py:465)
            base = name[: -len("weight")]
            scale = P[base + "weight"] / torch.sqrt(P[base + "running_var"] + 1e-5)
            shift = P[base + "bias"] - P[base + "running_mean"] * scale
            prep.w[base + "fold"] = (scale.float().contiguous(), shift.float().contiguous())
        if name.endswith("attention.query_proj.linear.weight"):
            base = name[: -len("query_proj.linear.weight")]
            cat = torch.cat([P[base + "query_proj.linear.weight"], P[base + "key_proj.linear.weight"],
                             P[base + "value_proj.linear.weight"]], 0).float().contiguous()
            prep.w[base + "qkv"] = cat
            _cat_planes(prep, base + "qkv", cat, name, tc_decoder)


def _relpos_attention(P, a, qkv, pos_proj, n_head):
    """RelativeMultiHeadAttention.forward (conformer.py:397-431), FP32, scores materialised like the reference
    ([B, H, T, T] content and positional scores, then the shifted sum, softmax WITHOUT padding mask, P.V)."""
    B, T, C3 = qkv.shape
    C = C3 // 3
    dh = C // n_head
    Z = B * n_head
    dev = qkv.device
    st = _stream()
    q = qkv[:, :, :C]
    # q + u_bias / q + v_bias (broadcast over batch and time): [B,T,C]
    qu = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    qv = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    qc = q.contiguous()
    capi.call("ctts_add_row_broadcast", qc, P[a + "attention.u_bias"], 1, B * T, C, qu, st)
    capi.call("ctts_add_row_broadcast", qc, P[a + "attention.v_bias"], 1, B * T, C, qv, st)
    content = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
    pscore = torch.empty(Z, T, T, device=dev, dtype=torch.float32)
    # content[z,t,s] = (q+u)[b,t,h,:] . k[b,s,h,:]      (k = qkv[..., C:2C])
    capi.call("ctts_batched_gemm_fp32", qu, qkv[:, :, C:], 1.0, None, 1, Z, n_head, T, dh, T,
              T * C, dh, C, T * C3, dh, C3, n_head * T * T, T * T, T, content, st)
    # pos[z,t,j] = (q+v)[b,t,h,:] . pos_proj[j,h,:]      (shared by all batch elements)
    capi.call("ctts_batched_gemm_fp32", qv, pos_proj, 1.0, None, 1, Z, n_head, T, dh, T,
              T * C, dh, C, 0, dh, C, n_head * T * T, T * T, T, pscore, st)
    ldp = (T + 15) // 16 * 16
    prob = torch.empty(Z, T, ldp, device=dev, dtype=torch.float32)
    capi.call("ctts_relshift_softmax", content, pscore, Z, T, ldp, math.sqrt(C), prob, st)
    vt = torch.empty(Z, dh, ldp, device=dev, dtype=torch.float32)
    capi.call("ctts_transpose_heads", qkv, B, T, C3, 2 * C, n_head, dh, ldp, vt, st)
    ctx = torch.empty(B, T, C, device=dev, dtype=torch.float32)
    capi.call("ctts_batched_gemm_fp32", prob, vt, 1.0, None, 1, Z, n_head, T, ldp, dh,
              n_head * T * ldp, T * ldp, ldp, n_head * dh * ldp, dh * ldp, ldp, T * C, dh, C, ctx, st)
    return ctx


def _relpos_attention_tc(P, a, qkv, qkv_planes, pos_proj, n_head):
    """The same attention on tcgen05 (bf16 planes): the 32-wide heads are zero padded to 64 columns so that one head is one
    k-block of the GEMM engine; content and positional scores are batched plane GEMMs, the shifted sum + softmax writes the
    probability PLANES directly (no fp32 probabilities in HBM), P.V is a third batched GEMM.  The [B*H, T, T] score tensors
    are still materialised in fp32, like the reference does."""
    B, T, C3 = qkv.shape
    C = C3 // 3
    H = n_head
    dh = C // H
    Z = B * H
    n = qkv_planes.n
    dev = qkv.device
    st = _stream()
    DHp = 64
    Cp = H * DHp
    Tp = (T + 7) // 8 * 8

    def padded(src, bias, rows, ld, c0):
        out = Planes.empty((rows, Cp), dev, n)
        capi.call("ctts_pad_heads_planes", src, bias, rows, ld, c0, H, dh, DHp, n, capi.ptr_array(out.p), st)
        return out

    qu = padded(qkv, P[a + "attention.u_bias"], B * T, C3, 0)
    qv = padded(qkv, P[a + "attention.v_bias"], B * T, C3, 0)
    kp = padded(qkv, None, B * T, C3, C)
    pp = padded(pos_proj, None, T, C, 0)
    act_view = (Cp, T, B, Cp, T * Cp)
    content = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
    pscore = torch.empty(Z, T, Tp, device=dev, dtype=torch.float32)
    gemm_batched_planes(qu, act_view, kp, act_view, (H, H, 0, DHp, H, 0, DHp, 1, Tp), H * T * Tp, T * Tp, 1.0, Z, T, DHp, Tp,
                        y=content)
    gemm_batched_planes(qv, act_view, pp, (Cp, T, 1, Cp, T * Cp), (H, H, 0, DHp, 0x7fffffff, 0, DHp, 1, Tp), H * T * Tp, T * Tp,
                        1.0, Z, T, DHp, Tp, y=pscore)
    prob = Planes.empty((Z, T, Tp), dev, n)
    capi.call("ctts_relshift_softmax_planes", content, pscore, Z, T, Tp, Tp, math.sqrt(C), n, capi.ptr_array(prob.p), st)
    vt = Planes.empty((Z, dh, Tp), dev, n)
    capi.call("ctts_transpose_v_planes", n, capi.ptr_array(qkv_planes.p), B, T, C, H, capi.ptr_array(vt.p), st)
    ctx_planes = Planes.empty((B, T, C), dev, n)
    gemm_batched_planes(prob, (T, T, Z, Tp, T * Tp), vt, (T, dh, Z, Tp, dh * Tp), (H, 1, 0, 0, 1, 0, 0, 1, C), T * C, dh, 1.0,
                        Z, T, T, dh, y_planes=ctx_planes)
    return ctx_planes


def _conformer_ffn(prep, P, p, x, n):
    """FeedForwardModule + half-step residual: LN -> Linear -> Swish -> Linear, * 0.5 + x (conformer.py:205-213,264-295)."""
    if n:
        tag = "#planes" if n == 2 else "#planes3"
        _, hp = layernorm_planes(x, P[p + "0.weight"], P[p + "0.bias"], 1e-5, n=n)
        _, gp = gemm_tc(hp, prep.w[p + "1.linear.weight" + tag], P[p + "1.linear.bias"], act=capi.ACT_SWISH,
                        want_fp32=False, want_planes=True)
        y, _ = gemm_tc(gp, prep.w[p + "4.linear.weight" + tag], P[p + "4.linear.bias"], alpha=0.5, residual=x)
        return y
    h = layernorm(x, P[p + "0.weight"], P[p + "0.bias"], 1e-5)
    g = conv_gemm(h, P[p + "1.linear.weight"], P[p + "1.linear.bias"], act=capi.ACT_SWISH)
    return conv_gemm(g, P[p + "4.linear.weight"], P[p + "4.linear.bias"], alpha=0.5, residual=x)


def _stack_conformer(prep, P, cfg, pre, x, lens, n_layers, n_head, kernel, n):
    B, T, C = x.shape
    W = prep.w
    st = _stream()
    tc = n > 0
    tag = "#planes" if n == 2 else "#planes3"
    for i in range(n_layers):
        lp = "%slayer_stack.%d.sequential." % (pre, i)
        x = _conformer_ffn(prep, P, lp + "0.module.sequential.", x, n)
        a = lp + "1.module."
        pos = abs_table(prep, P, a + "positional_encoding", T, C, cfg["max_seq_len"], x.device)[:T].contiguous()
        pos_proj = conv_gemm(pos.view(1, T, C), P[a + "attention.pos_proj.linear.weight"]).view(T, C)
        m = lp + "2.module.sequential."
        if tc:
            _, hp = layernorm_planes(x, P[a + "layer_norm.weight"], P[a + "layer_norm.bias"], 1e-5, n=n)
            qkv, qkvp = gemm_tc(hp, W[a + "attention.qkv" + tag], want_planes=True)
            ctxp = _relpos_attention_tc(P, a, qkv, qkvp, pos_proj, n_head)
            x, _ = gemm_tc(ctxp, W[a + "attention.out_proj.linear.weight" + tag], residual=x)
            _, hp = layernorm_planes(x, P[m + "0.weight"], P[m + "0.bias"], 1e-5, n=n)
            pw, _ = gemm_tc(hp, W[m + "2.conv.weight" + tag], P[m + "2.conv.bias"])
        else:
            h = layernorm(x, P[a + "layer_norm.weight"], P[a + "layer_norm.bias"], 1e-5)
            qkv = conv_gemm(h, W[a + "attention.qkv"])
            ctx = _relpos_attention(P, a, qkv, pos_proj, n_head)
            x = conv_gemm(ctx, P[a + "attention.out_proj.linear.weight"], residual=x)
            h = layernorm(x, P[m + "0.weight"], P[m + "0.bias"], 1e-5)
            pw = conv_gemm(h, W[m + "2.conv.weight"], P[m + "2.conv.bias"])
        g = torch.empty(B, T, C, device=x.device, dtype=torch.float32)
        capi.call("ctts_glu", pw, B * T, C, g, st)
        d = torch.empty_like(g)
        fold = W[m + "5.fold"]
        capi.call("ctts_dwconv_bn_swish", g, P[m + "4.conv.weight"], kernel, fold[0], fold[1], B, T, C, d, st)
        if tc:
            x, _ = gemm_tc(split_planes(d, n), W[m + "7.conv.weight" + tag], P[m + "7.conv.bias"], residual=x)
        else:
            x = conv_gemm(d, W[m + "7.conv.weight"], P[m + "7.conv.bias"], residual=x)
        x = _conformer_ffn(prep, P, lp + "3.module.sequential.", x, n)
        x = layernorm(x, P[lp + "4.weight"], P[lp + "4.bias"], 1e-5, lens)
    return x


def encoder_conformer(prep, P, cfg, tokens, src_lens):
    c = cfg["conformer"]
    x, word = embed_abs(prep, P, cfg, c["encoder_hidden"], tokens)
    return _stack_conformer(prep, P, cfg, "encoder.", x, src_lens, c["encoder_layer"], c["encoder_head"],
                            c["conv_kernel_size"], _n_planes(prep, "encoder.", False)), word


def decoder_conformer(prep, P, cfg, x, mel_lens, math_mode):
    c = cfg["conformer"]
    x = add_abs_positions(prep, P, cfg, "decoder.position_enc", x)
    return _stack_conformer(prep, P, cfg, "decoder.", x, mel_lens, c["decoder_layer"], c["decoder_head"],
                            c["conv_kernel_size"], _n_planes(prep, "decoder.", math_mode == "bf16x3")), None


ENCODERS = {"transformer": encoder_transformer, "fastformer": encoder_fastformer, "conformer": encoder_conformer}
DECODERS = {"transformer": decoder_transformer, "fastformer": decoder_fastformer, "conformer": decoder_conformer}
PREPARE = {"transformer": prepare_transformer, "fastformer": prepare_fastformer, "conformer": prepare_conformer}
