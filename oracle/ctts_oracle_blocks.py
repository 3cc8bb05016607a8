"""CPU oracle, block types other than transformer_fs2.  TEST INFRASTRUCTURE -- NOT PRODUCT (see ctts_oracle.py).

Functional fp32 PyTorch restatements of model/transformers/{transformer,fastformer,conformer}.py (dropout = identity;
BatchNorm on running statistics in eval mode, on batch statistics with training=True), over the reference's state_dict
key names.
"""
import math

import torch
import torch.nn.functional as F

from .ctts_oracle import _batch_norm_train, sinusoid_table_interleaved


def _no_truncation(training, T, cfg):
    """In training mode the reference truncates sequences longer than max_seq_len (transformer.py:128-143 and the same
    branch in fastformer.py / conformer.py); the training oracle restates the un-truncated case only."""
    if training and T > cfg["max_seq_len"]:
        raise NotImplementedError("training-mode oracle: sequence length %d > max_seq_len %d (the reference truncates)"
                                  % (T, cfg["max_seq_len"]))


def _abs_positions(P, key, T, d_model, max_seq_len):
    """`position_enc[:, :T]`, or a freshly generated table when T > max_seq_len in eval mode
    (transformer.py:65-74,128-145; fastformer.py:54-64,103-122; conformer.py:75-84,140-154,331-339)."""
    if T > max_seq_len:
        return sinusoid_table_interleaved(T, d_model)[:T]
    return P[key][0, :T]


# ---------------------------------------------------------------------------------------------
# "transformer": post-LN FFT block (model/transformers/transformer.py)
# ---------------------------------------------------------------------------------------------
def _mha_transformer(P, pre, x, pad_mask, n_head):
    """MultiHeadAttention + ScaledDotProductAttention, transformer.py:181-252."""
    B, T, C = x.shape
    dk = C // n_head
    q = F.linear(x, P[pre + "w_qs.linear.weight"]).view(B, T, n_head, dk).transpose(1, 2)
    k = F.linear(x, P[pre + "w_ks.linear.weight"]).view(B, T, n_head, dk).transpose(1, 2)
    v = F.linear(x, P[pre + "w_vs.linear.weight"]).view(B, T, n_head, dk).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dk)
    s = s.masked_fill(pad_mask[:, None, None, :], float("-inf"))
    o = torch.matmul(torch.softmax(s, -1), v).transpose(1, 2).reshape(B, T, C)
    o = F.linear(o, P[pre + "fc.linear.weight"])
    return F.layer_norm(o + x, (C,), P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-5)


def _ffn_transformer(P, pre, x, kernel):
    """PositionwiseFeedForward, transformer.py:255-288."""
    C = x.shape[-1]
    h = F.conv1d(x.transpose(1, 2), P[pre + "w_1.weight"], P[pre + "w_1.bias"], padding=(kernel[0] - 1) // 2)
    h = F.conv1d(F.relu(h), P[pre + "w_2.weight"], P[pre + "w_2.bias"], padding=(kernel[1] - 1) // 2).transpose(1, 2)
    return F.layer_norm(h + x, (C,), P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-5)


def _stack_transformer(P, pre, x, pad_mask, n_layers, n_head, kernel, taps=None):
    keep = (~pad_mask)[:, :, None]
    for i in range(n_layers):
        lp = "%slayer_stack.%d." % (pre, i)
        x = _mha_transformer(P, lp + "slf_attn.", x, pad_mask, n_head) * keep
        x = _ffn_transformer(P, lp + "pos_ffn.", x, kernel) * keep
        if taps is not None:
            taps["%slayer_stack.%d" % (pre, i)] = x
    return x


def encoder_transformer(P, cfg, tokens, pad_mask, taps=None, training=False, stats_out=None):
    """TextEncoder.forward, transformer.py:56-83."""
    c = cfg["transformer"]
    _no_truncation(training, tokens.shape[1], cfg)
    word = F.embedding(tokens, P["encoder.src_word_emb.weight"], padding_idx=0)
    x = word + _abs_positions(P, "encoder.position_enc", tokens.shape[1], c["encoder_hidden"], cfg["max_seq_len"])[None]
    x = _stack_transformer(P, "encoder.", x, pad_mask, c["encoder_layer"], c["encoder_head"], c["conv_kernel_size"], taps)
    return x, word


def decoder_transformer(P, cfg, x, pad_mask, taps=None, training=False, stats_out=None):
    """Decoder.forward, transformer.py:121-154 (eval: no truncation when T > max_seq_len; training: see
    _no_truncation)."""
    c = cfg["transformer"]
    T = x.shape[1]
    _no_truncation(training, T, cfg)
    if T <= cfg["max_seq_len"]:
        x = x + P["decoder.position_enc"][0, :T][None]
    else:
        x = x + sinusoid_table_interleaved(T, c["decoder_hidden"])[None]
    x = _stack_transformer(P, "decoder.", x, pad_mask, c["decoder_layer"], c["decoder_head"], c["conv_kernel_size"], taps)
    return x, pad_mask


# ---------------------------------------------------------------------------------------------
# "fastformer": additive attention (model/transformers/fastformer.py) -- quirks of SURVEY.md section 7 item 1
# ---------------------------------------------------------------------------------------------
def _fast_attention(P, pre, logit_pre, h, pad_mask, n_heads_eff):
    """FastAttention.forward, fastformer.py:296-345.  `n_heads_eff` = the ctor's dim_head (128 at d 256)."""
    B, T, C = h.shape
    hs = C // n_heads_eff
    add = ((1.0 - pad_mask.float()) * -10000.0)[:, None, :]     # inverted mask: VALID positions get -10000
    q = F.linear(h, P[pre + "query.weight"], P[pre + "query.bias"])
    k = F.linear(h, P[pre + "key.weight"], P[pre + "key.bias"])
    qs = F.linear(q, P[logit_pre + "to_q_attn_logits.weight"], P[logit_pre + "to_q_attn_logits.bias"]).transpose(1, 2) \
        / hs ** 0.5
    qs = qs + add
    qw = torch.softmax(qs, -1).unsqueeze(2)                       # [B, H, 1, T]
    ql = q.view(B, T, n_heads_eff, hs).permute(0, 2, 1, 3)        # [B, H, T, hs]
    pooled_q = torch.matmul(qw, ql).transpose(1, 2).reshape(B, 1, C)
    qk = k * pooled_q
    ks = (F.linear(qk, P[logit_pre + "to_k_attn_logits.weight"], P[logit_pre + "to_k_attn_logits.bias"])
          / hs ** 0.5).transpose(1, 2)
    ks = ks + add
    kw = torch.softmax(ks, -1).unsqueeze(2)
    kl = qk.view(B, T, n_heads_eff, hs).permute(0, 2, 1, 3)
    pooled_k = torch.matmul(kw, kl)                               # [B, H, 1, hs]
    wv = (pooled_k * ql).transpose(1, 2).reshape(B, T, C)
    return F.linear(wv, P[pre + "transform.weight"], P[pre + "transform.bias"]) + q


def _stack_fastformer(P, pre, x, pad_mask, n_layers, d_head, kernel, taps=None):
    """FFTBlock.forward, fastformer.py:163-171; to_{q,k}_attn_logits are tied to layer 0 (:157-161)."""
    keep = (~pad_mask)[:, :, None]
    C = x.shape[-1]
    tied = "%slayer_stack.layers.0.0.fn." % pre
    for i in range(n_layers):
        a = "%slayer_stack.layers.%d.0." % (pre, i)
        f = "%slayer_stack.layers.%d.1." % (pre, i)
        h = F.layer_norm(x, (C,), P[a + "norm.weight"], P[a + "norm.bias"], 1e-5)
        x = (_fast_attention(P, a + "fn.", tied, h, pad_mask, d_head) + x) * keep
        h = F.layer_norm(x, (C,), P[f + "norm.weight"], P[f + "norm.bias"], 1e-5)
        g = F.conv1d(h.transpose(1, 2), P[f + "fn.w_1.weight"], P[f + "fn.w_1.bias"], padding=(kernel[0] - 1) // 2)
        g = F.conv1d(F.gelu(g), P[f + "fn.w_2.weight"], P[f + "fn.w_2.bias"], padding=(kernel[1] - 1) // 2)
        x = (g.transpose(1, 2) + x) * keep
        if taps is not None:
            taps["%slayer_stack.%d" % (pre, i)] = x
    return x


def encoder_fastformer(P, cfg, tokens, pad_mask, taps=None, training=False, stats_out=None):
    c = cfg["transformer"]  # fastformer reads the `transformer` section, fastformer.py:24-34
    _no_truncation(training, tokens.shape[1], cfg)
    word = F.embedding(tokens, P["encoder.src_word_emb.weight"], padding_idx=0)
    x = word + _abs_positions(P, "encoder.position_enc", tokens.shape[1], c["encoder_hidden"], cfg["max_seq_len"])[None]
    d_head = c["encoder_hidden"] // c["encoder_head"]
    return _stack_fastformer(P, "encoder.", x, pad_mask, c["encoder_layer"], d_head, c["conv_kernel_size"], taps), word


def decoder_fastformer(P, cfg, x, pad_mask, taps=None, training=False, stats_out=None):
    c = cfg["transformer"]
    T = x.shape[1]
    _no_truncation(training, T, cfg)
    x = x + _abs_positions(P, "decoder.position_enc", T, c["decoder_hidden"], cfg["max_seq_len"])[None]
    d_head = c["decoder_hidden"] // c["decoder_head"]
    return _stack_fastformer(P, "decoder.", x, pad_mask, c["decoder_layer"], d_head, c["conv_kernel_size"], taps), pad_mask


# ---------------------------------------------------------------------------------------------
# "conformer" (model/transformers/conformer.py) -- quirks of SURVEY.md section 7 item 2
# ---------------------------------------------------------------------------------------------
def _rel_shift(s):
    """RelativeMultiHeadAttention._relative_shift, conformer.py:423-431."""
    B, H, T1, T2 = s.shape
    z = s.new_zeros(B, H, T1, 1)
    p = torch.cat([z, s], dim=-1).view(B, H, T2 + 1, T1)
    return p[:, :, 1:].reshape(B, H, T1, T2)


def _conformer_block(P, pre, x, pos, n_head, kernel, training=False, stats_out=None):
    """ConformerBlock.sequential, conformer.py:205-246 (no attention mask is passed: :242-246).  training: the conv
    module's BatchNorm1d uses batch statistics over all B x T positions (padded ones included); dropout = identity."""
    B, T, C = x.shape
    dh = C // n_head

    def ffn(p, v):
        h = F.layer_norm(v, (C,), P[p + "0.weight"], P[p + "0.bias"], 1e-5)
        h = F.linear(h, P[p + "1.linear.weight"], P[p + "1.linear.bias"])
        h = h * torch.sigmoid(h)
        return F.linear(h, P[p + "4.linear.weight"], P[p + "4.linear.bias"])

    x = ffn(pre + "sequential.0.module.sequential.", x) * 0.5 + x
    a = pre + "sequential.1.module."
    h = F.layer_norm(x, (C,), P[a + "layer_norm.weight"], P[a + "layer_norm.bias"], 1e-5)
    q = F.linear(h, P[a + "attention.query_proj.linear.weight"]).view(B, T, n_head, dh)
    k = F.linear(h, P[a + "attention.key_proj.linear.weight"]).view(B, T, n_head, dh).permute(0, 2, 1, 3)
    v = F.linear(h, P[a + "attention.value_proj.linear.weight"]).view(B, T, n_head, dh).permute(0, 2, 1, 3)
    pe = F.linear(pos, P[a + "attention.pos_proj.linear.weight"]).view(1, T, n_head, dh).expand(B, -1, -1, -1)
    content = torch.matmul((q + P[a + "attention.u_bias"]).transpose(1, 2), k.transpose(2, 3))
    pscore = torch.matmul((q + P[a + "attention.v_bias"]).transpose(1, 2), pe.permute(0, 2, 3, 1))
    score = (content + _rel_shift(pscore)) / math.sqrt(C)
    ctx = torch.matmul(torch.softmax(score, -1), v).transpose(1, 2).reshape(B, T, C)
    x = F.linear(ctx, P[a + "attention.out_proj.linear.weight"]) + x
    c = pre + "sequential.2.module.sequential."
    h = F.layer_norm(x, (C,), P[c + "0.weight"], P[c + "0.bias"], 1e-5).transpose(1, 2)
    h = F.conv1d(h, P[c + "2.conv.weight"], P[c + "2.conv.bias"])
    o, g = h.chunk(2, dim=1)
    h = o * torch.sigmoid(g)
    h = F.conv1d(h, P[c + "4.conv.weight"], None, padding=(kernel - 1) // 2, groups=C)
    if training:
        h = _batch_norm_train(h, P, c + "5.", stats_out)
    else:
        h = F.batch_norm(h, P[c + "5.running_mean"], P[c + "5.running_var"], P[c + "5.weight"], P[c + "5.bias"], False,
                         0.1, 1e-5)
    h = h * torch.sigmoid(h)
    h = F.conv1d(h, P[c + "7.conv.weight"], P[c + "7.conv.bias"]).transpose(1, 2)
    x = h + x
    x = ffn(pre + "sequential.3.module.sequential.", x) * 0.5 + x
    return F.layer_norm(x, (C,), P[pre + "sequential.4.weight"], P[pre + "sequential.4.bias"], 1e-5)


def _stack_conformer(P, pre, x, pad_mask, n_layers, n_head, kernel, d_model, max_seq_len, taps=None, training=False,
                     stats_out=None):
    keep = (~pad_mask)[:, :, None]
    T = x.shape[1]
    for i in range(n_layers):
        lp = "%slayer_stack.%d." % (pre, i)
        pos = _abs_positions(P, lp + "sequential.1.module.positional_encoding", T, d_model, max_seq_len)[None]
        x = _conformer_block(P, lp, x, pos, n_head, kernel, training, stats_out) * keep
        if taps is not None:
            taps["%slayer_stack.%d" % (pre, i)] = x
    return x


def encoder_conformer(P, cfg, tokens, pad_mask, taps=None, training=False, stats_out=None):
    c = cfg["conformer"]
    _no_truncation(training, tokens.shape[1], cfg)
    word = F.embedding(tokens, P["encoder.src_word_emb.weight"], padding_idx=0)
    x = word + _abs_positions(P, "encoder.position_enc", tokens.shape[1], c["encoder_hidden"], cfg["max_seq_len"])[None]
    x = _stack_conformer(P, "encoder.", x, pad_mask, c["encoder_layer"], c["encoder_head"], c["conv_kernel_size"],
                         c["encoder_hidden"], cfg["max_seq_len"], taps, training, stats_out)
    return x, word


def decoder_conformer(P, cfg, x, pad_mask, taps=None, training=False, stats_out=None):
    c = cfg["conformer"]
    T = x.shape[1]
    _no_truncation(training, T, cfg)
    x = x + _abs_positions(P, "decoder.position_enc", T, c["decoder_hidden"], cfg["max_seq_len"])[None]
    x = _stack_conformer(P, "decoder.", x, pad_mask, c["decoder_layer"], c["decoder_head"], c["conv_kernel_size"],
                         c["decoder_hidden"], cfg["max_seq_len"], taps, training, stats_out)
    return x, pad_mask


ENCODERS = {"transformer": encoder_transformer, "fastformer": encoder_fastformer, "conformer": encoder_conformer}
DECODERS = {"transformer": decoder_transformer, "fastformer": decoder_fastformer, "conformer": decoder_conformer}
