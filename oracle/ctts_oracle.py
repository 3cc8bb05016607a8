"""CPU oracle for the CompTransTTS acoustic-model forward path.  TEST INFRASTRUCTURE -- NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs (`cpu_baseline`,
`--impl reference`) may import this file.  The product path (`comprehensive-transformer-tts_b200/`)
never does, and fails loudly when its CUDA library is missing.

What it is: a plain fp32 PyTorch-CPU restatement, written as pure functions over a flat
state_dict (`name -> tensor`, the reference's own key names), of what the
reference computes in `CompTransTTS.forward` (model/CompTransTTS.py:64-152).  The numerics of
the individual ops (conv1d, linear, layer_norm, softmax, gelu, bucketize) are torch's, exactly as
in the reference, whose arithmetic also lives in torch (pinned torch==1.7.0,
requirements.txt:24; this image has torch 2.11).  Each function cites the reference lines it
follows.

Parity pin: the reference has no tests and no golden vectors.  The pin is
therefore created here: `tests/golden/make_golden.py` imports the UNMODIFIED reference from
`/root/reference` in the build container, loads the same synthetic state_dict into it and into
this oracle, and commits the reference's outputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against those fixtures (CPU, no reference
needed), and `tests/test_oracle_vs_reference.py` re-checks against the live reference whenever
`/root/reference` exists.

Scope: block types transformer_fs2 / transformer / fastformer / conformer (the latter three in
ctts_oracle_blocks.py); prosody "none" and "liu2021" (eval mode: predictors; training mode: the
reference encoders); inference (free-running), supervised teacher-forced, and unsupervised
(aligner + MAS) branches of the VarianceAdaptor.  Eval mode: dropout is identity, BatchNorm uses
running statistics.  `training=True` restates model.train() with every dropout probability 0
(BatchNorm on batch statistics); it is differentiable, and torch.autograd through it is the oracle
for the backward pass -- pinned the same way by tests/golden/make_golden_train.py /
tests/test_oracle_train.py (outputs, gradients of a fixed objective w.r.t. every parameter,
BatchNorm buffers after the step).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# helpers (utils/tools.py)
# ----------------------------------------------------------------------------------------------
def pad_mask_from_lengths(lengths, max_len=None):
    """True = padding.  utils/tools.py:188-196."""
    if max_len is None:
        max_len = int(lengths.max().item())
    ids = torch.arange(max_len, device=lengths.device)[None, :]
    return ids >= lengths[:, None]


def positions_of(flag_src, padding_idx=0):
    """utils/tools.py:640-652 -- running count of non-pad entries, pads keep `padding_idx`."""
    keep = flag_src.ne(padding_idx).int()
    return (torch.cumsum(keep, dim=1).type_as(keep) * keep).long() + padding_idx


def sinusoid_table_fs2(n_rows, dim, padding_idx=0):
    """fairseq-style [sin | cos] table.  model/transformers/blocks.py:66-83."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float) * -step)
    ang = torch.arange(n_rows, dtype=torch.float).unsqueeze(1) * freq.unsqueeze(0)
    tab = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1).view(n_rows, -1)
    if dim % 2 == 1:
        tab = torch.cat([tab, torch.zeros(n_rows, 1)], dim=1)
    tab[padding_idx, :] = 0
    return tab


def sinusoid_table_interleaved(n_position, d_hid):
    """Interleaved sin/cos table computed in float64 then cast.  blocks.py:26-46."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.FloatTensor(tab)


def fs2_positional(flag_src, dim, init_size):
    """SinusoidalPositionalEmbedding.forward, blocks.py:85-104: table grows to seq_len+1 if needed."""
    seq_len = flag_src.shape[1]
    tab = sinusoid_table_fs2(max(init_size, seq_len + 1), dim, 0)
    pos = positions_of(flag_src, 0)
    return tab.index_select(0, pos.reshape(-1)).view(flag_src.shape[0], seq_len, dim)


def durations_to_mel2ph(dur, dur_padding=None):
    """utils/tools.py:598-628 with alpha = 1: frame -> 1-based phoneme index, 0 beyond sum(dur)."""
    dur = torch.round(dur.float()).long()
    if dur_padding is not None:
        dur = dur * (1 - dur_padding.long())
    csum = torch.cumsum(dur, 1)
    prev = F.pad(csum, [1, -1])
    total = int(dur.sum(-1).max().item())
    t = torch.arange(total)[None, None]
    hit = (t >= prev[:, :, None]) & (t < csum[:, :, None])
    idx = torch.arange(1, dur.shape[1] + 1)[None, :, None]
    return (idx * hit.long()).sum(1)


def length_regulate(x, duration, max_len=None):
    """LengthRegulator.LR/expand (model/modules.py:1222-1249) + pad (utils/tools.py:577-595).

    Row j of utterance b is repeated max(int(duration[b, j]), 0) times (truncation, not
    rounding), rows are zero padded to `max_len` (or the batch maximum).  Returns the expanded
    tensor and the per-utterance lengths (before padding).
    """
    reps = duration.to(torch.float64).trunc().clamp(min=0).long() if duration.is_floating_point() \
        else duration.clamp(min=0).long()
    lens = reps.sum(1)
    out_len = int(max_len) if max_len else int(lens.max().item())
    B, S = reps.shape
    out = x.new_zeros(B, out_len, x.shape[-1])
    for b in range(B):
        idx = torch.repeat_interleave(torch.arange(S), reps[b])
        n = idx.numel()
        n = min(n, out_len)   # F.pad with a negative amount crops (utils/tools.py:586-592); mel_len keeps the full length
        out[b, :n] = x[b, idx[:n]]
    return out, lens


# ----------------------------------------------------------------------------------------------
# transformer_fs2 blocks (model/transformers/transformer_fs2.py)
# ----------------------------------------------------------------------------------------------
def _mha_fs2(P, pre, x, pad_mask, n_head):
    """EncSALayer's self-attention: F.multi_head_attention_forward with in_proj_weight [3C, C],
    no biases, key_padding_mask (transformer_fs2.py:385-394).  x: [B, T, C]."""
    B, T, C = x.shape
    dh = C // n_head
    qkv = F.linear(x, P[pre + "self_attn.in_proj_weight"])
    q, k, v = qkv.split(C, dim=-1)
    q = q.view(B, T, n_head, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, T, n_head, dh).transpose(1, 2)
    v = v.view(B, T, n_head, dh).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2))
    s = s.masked_fill(pad_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(B, T, C)
    return F.linear(o, P[pre + "self_attn.out_proj.weight"])


def _ffn_fs2(P, pre, x, kernel_size, act="gelu"):
    """TransformerFFNLayer.forward, transformer_fs2.py:220-239 (padding SAME)."""
    h = F.conv1d(x.transpose(1, 2), P[pre + "ffn.ffn_1.weight"], P[pre + "ffn.ffn_1.bias"],
                 padding=kernel_size // 2).transpose(1, 2)
    h = h * kernel_size ** -0.5
    if act == "gelu":
        h = F.gelu(h)
    elif act == "relu":
        h = F.relu(h)
    else:
        raise NotImplementedError(act)
    return F.linear(h, P[pre + "ffn.ffn_2.weight"], P[pre + "ffn.ffn_2.bias"])


def fft_blocks_fs2(P, pre, x, pad_mask, n_layers, n_head, kernel_size, use_pos, init_size, act="gelu",
                   taps=None):
    """FFTBlocks.forward (transformer_fs2.py:47-72) + EncSALayer.forward (:176-200). x: [B,T,C]."""
    keep = (~pad_mask).float()[:, :, None]
    C = x.shape[-1]
    if use_pos:
        x = x + P[pre + "pos_embed_alpha"] * fs2_positional(x[..., 0], C, init_size)
    x = x * keep
    for i in range(n_layers):
        lp = "%slayers.%d.op." % (pre, i)
        h = F.layer_norm(x, (C,), P[lp + "layer_norm1.weight"], P[lp + "layer_norm1.bias"], 1e-12)
        x = (x + _mha_fs2(P, lp, h, pad_mask, n_head)) * keep
        h = F.layer_norm(x, (C,), P[lp + "layer_norm2.weight"], P[lp + "layer_norm2.bias"], 1e-12)
        x = (x + _ffn_fs2(P, lp, h, kernel_size, act)) * keep
        if taps is not None:
            taps["%slayers.%d" % (pre, i)] = x
    x = F.layer_norm(x, (C,), P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-5) * keep
    return x


def encoder_fs2(P, cfg, tokens, pad_mask, taps=None):
    """TextEncoder.forward / forward_embedding, transformer_fs2.py:100-119."""
    c = cfg["transformer_fs2"]
    C = c["encoder_hidden"]
    word = math.sqrt(C) * F.embedding(tokens, P["encoder.embed_tokens.weight"], padding_idx=0)
    x = word + fs2_positional(tokens, C, cfg["max_seq_len"])
    x = fft_blocks_fs2(P, "encoder.", x, pad_mask, c["encoder_layer"], c["encoder_head"], c["ffn_kernel_size"],
                       False, 0, cfg["variance_predictor"]["ffn_act"], taps)
    return x, word


def decoder_fs2(P, cfg, x, pad_mask, taps=None):
    """Decoder (FFTBlocks with positional embedding, init_size 2*max_seq_len), transformer_fs2.py:122-134."""
    c = cfg["transformer_fs2"]
    return fft_blocks_fs2(P, "decoder.", x, pad_mask, c["decoder_layer"], c["decoder_head"], c["ffn_kernel_size"],
                          True, cfg["max_seq_len"] * 2, cfg["variance_predictor"]["ffn_act"], taps), pad_mask


# ----------------------------------------------------------------------------------------------
# variance adaptor (model/modules.py)
# ----------------------------------------------------------------------------------------------
def _conv_relu_ln_stack(P, pre, xs, n_layers, kernel, pad_mask=None):
    """The [ConstantPad1d, Conv1d, ReLU, LayerNorm(dim=1, eps 1e-12), Dropout] stacks of
    DurationPredictor (modules.py:1277-1288, 1299-1304) and PitchPredictor (:1330-1338, 1351-1352).
    xs: [B, T, C] (kept token-major here; the reference transposes to [B, C, T])."""
    for l in range(n_layers):
        w = P["%sconv.%d.1.weight" % (pre, l)]
        h = F.conv1d(xs.transpose(1, 2), w, P["%sconv.%d.1.bias" % (pre, l)], padding=(kernel - 1) // 2)
        h = F.relu(h).transpose(1, 2)
        xs = F.layer_norm(h, (h.shape[-1],), P["%sconv.%d.3.weight" % (pre, l)], P["%sconv.%d.3.bias" % (pre, l)],
                          1e-12)
        if pad_mask is not None:
            xs = xs * (~pad_mask).float()[:, :, None]
    return xs


def duration_predictor(P, cfg, x, src_mask):
    """DurationPredictor.forward (dur_loss 'mse'), modules.py:1299-1310 -> log durations [B, S]."""
    vp = cfg["variance_predictor"]
    pre = "variance_adaptor.duration_predictor."
    h = _conv_relu_ln_stack(P, pre, x, vp["dur_predictor_layers"], vp["dur_predictor_kernel"], src_mask)
    out = F.linear(h, P[pre + "linear.weight"], P[pre + "linear.bias"])
    return (out * (~src_mask).float()[:, :, None]).squeeze(-1)


def pitch_style_predictor(P, cfg, pre, xs):
    """PitchPredictor.forward (also EnergyPredictor), modules.py:1343-1356.  xs: [B, T, idim]."""
    vp = cfg["variance_predictor"]
    xs = xs + P[pre + "pos_embed_alpha"] * fs2_positional(xs[..., 0], xs.shape[-1], 4096)
    h = _conv_relu_ln_stack(P, pre, xs, vp["predictor_layers"], vp["predictor_kernel"], None)
    return F.linear(h, P[pre + "linear.weight"], P[pre + "linear.bias"])


F0_BIN, F0_MAX, F0_MIN = 256, 1100.0, 50.0


def f0_to_coarse(f0):
    """utils/pitch_tools.py:20-36 (torch branch): mel-scale bucket index in [1, 255]."""
    mel_min = 1127 * np.log(1 + F0_MIN / 700)
    mel_max = 1127 * np.log(1 + F0_MAX / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    pos = f0_mel > 0
    f0_mel = torch.where(pos, (f0_mel - mel_min) * (F0_BIN - 2) / (mel_max - mel_min) + 1, f0_mel)
    f0_mel = torch.where(f0_mel <= 1, torch.ones_like(f0_mel), f0_mel)
    f0_mel = torch.where(f0_mel > F0_BIN - 1, torch.full_like(f0_mel, F0_BIN - 1), f0_mel)
    return (f0_mel + 0.5).long()


def cwt_to_f0_norm(cwt_spec, mean, std, n_frames, pitch_cfg):
    """cwt2f0_norm -> cwt2f0 -> inverse_cwt_torch -> norm_f0, utils/pitch_tools.py:258-294,39-48.
    The standardisation runs over ALL (padded) frames with the unbiased std."""
    n_scales = cwt_spec.shape[-1]
    b = (torch.arange(0, n_scales).float()[None, None, :] + 1 + 2.5) ** (-2.5)
    rec = (cwt_spec * b).sum(-1)
    rec = (rec - rec.mean(-1, keepdim=True)) / rec.std(-1, keepdim=True)
    f0 = (rec * std[:, None] + mean[:, None]).exp()
    if n_frames > f0.shape[1]:
        f0 = torch.cat([f0] + [f0[:, -1:]] * (n_frames - f0.shape[1]), 1)
    if pitch_cfg["pitch_norm"] == "standard":
        f0 = (f0 - pitch_cfg["f0_mean"]) / pitch_cfg["f0_std"]
    if pitch_cfg["pitch_norm"] == "log":
        f0 = torch.log2(f0 + pitch_cfg["pitch_norm_eps"])
    return f0


def denorm_f0(f0, uv, pitch_cfg, pitch_padding=None):
    """utils/pitch_tools.py:69-82."""
    if pitch_cfg["pitch_norm"] == "standard":
        f0 = f0 * pitch_cfg["f0_std"] + pitch_cfg["f0_mean"]
    if pitch_cfg["pitch_norm"] == "log":
        f0 = 2 ** f0
    if uv is not None and pitch_cfg["use_uv"]:
        f0 = torch.where(uv > 0, torch.zeros_like(f0), f0)
    if pitch_padding is not None:
        f0 = torch.where(pitch_padding, torch.zeros_like(f0), f0)
    return f0


def pitch_embedding_cwt(P, cfg, pcfg, decoder_inp, f0, uv, mel2ph, control, x_org):
    """get_pitch_embedding, pitch_type == 'cwt' branch, modules.py:890-948."""
    pitch_cfg = pcfg["preprocessing"]["pitch"]
    pre = "variance_adaptor."
    g = cfg["variance_predictor"]["predictor_grad"]          # modules.py:904: gradient scaling, values unchanged
    decoder_inp = decoder_inp.detach() + g * (decoder_inp - decoder_inp.detach())
    h = F.linear(decoder_inp, P[pre + "cwt_predictor.0.weight"], P[pre + "cwt_predictor.0.bias"])
    cwt = pitch_style_predictor(P, cfg, pre + "cwt_predictor.1.", h) * control
    s = F.relu(F.linear(x_org[:, 0, :], P[pre + "cwt_stats_layers.0.weight"], P[pre + "cwt_stats_layers.0.bias"]))
    s = F.relu(F.linear(s, P[pre + "cwt_stats_layers.2.weight"], P[pre + "cwt_stats_layers.2.bias"]))
    stats = F.linear(s, P[pre + "cwt_stats_layers.4.weight"], P[pre + "cwt_stats_layers.4.bias"])
    f0_mean, f0_std = stats[:, 0], stats[:, 1]
    if f0 is None:
        std = f0_std * cfg["variance_predictor"]["cwt_std_scale"]
        f0 = cwt_to_f0_norm(cwt[:, :, :10], f0_mean, std, mel2ph.shape[1], pitch_cfg)
        if pitch_cfg["use_uv"]:
            uv = cwt[:, :, -1] > 0
    f0_denorm = denorm_f0(f0, uv, pitch_cfg, None)
    emb = F.embedding(f0_to_coarse(f0_denorm), P[pre + "pitch_embed.weight"], padding_idx=0)
    pred = {"pitch_pred": None, "f0_denorm": f0_denorm, "cwt": cwt, "f0_mean": f0_mean, "f0_std": f0_std}
    return pred, emb


def phoneme_level_pitch(text, src_len, mel2ph, mel_len, pitch_frame):
    """VarianceAdaptor.get_phoneme_level_pitch (modules.py:874-880) + utils/tools.py:47-53 + pad_1D: per utterance the
    mean of the frame-level f0 over the frames of each phoneme (scatter-mean by mel2ph), zero for phonemes without frames."""
    B = pitch_frame.shape[0]
    rows = []
    for b in range(B):
        s, m = int(src_len[b]), int(mel_len[b])
        idx = mel2ph[b, :m].long() - 1
        tot = torch.zeros(s).scatter_add(0, idx, pitch_frame[b, :m].float())
        num = torch.zeros(s).scatter_add(0, idx, torch.ones(m)).clamp_min(1)
        rows.append(tot / num)
    L = max(r.numel() for r in rows)
    out = torch.zeros(B, L)
    for b, r in enumerate(rows):
        out[b, : r.numel()] = r
    return out


def pitch_embedding_frame_or_ph(P, cfg, pcfg, decoder_inp, f0, uv, mel2ph, control, x_org):
    """get_pitch_embedding, pitch_type 'ph' and 'frame' (pitch_ar False) branches, modules.py:890-906,927-948."""
    pitch_cfg = pcfg["preprocessing"]["pitch"]
    pre = "variance_adaptor."
    g = cfg["variance_predictor"]["predictor_grad"]
    if pitch_cfg["pitch_type"] == "ph":
        inp = x_org.detach() + g * (x_org - x_org.detach())
        pitch_padding = x_org.sum().abs() == 0          # a SCALAR in the reference (modules.py:894): kept
        pred = pitch_style_predictor(P, cfg, pre + "pitch_predictor.", inp) * control
        if f0 is None:
            f0 = pred[:, :, 0]
        f0_denorm = denorm_f0(f0, None, pitch_cfg, pitch_padding=pitch_padding)
        pitch = F.pad(f0_to_coarse(f0_denorm), [1, 0])
        pitch = torch.gather(pitch, 1, mel2ph)
    else:
        inp = decoder_inp.detach() + g * (decoder_inp - decoder_inp.detach())
        pitch_padding = mel2ph == 0
        pred = pitch_style_predictor(P, cfg, pre + "pitch_predictor.", inp) * control
        if f0 is None:
            f0 = pred[:, :, 0]
        if pitch_cfg["use_uv"] and uv is None:
            uv = pred[:, :, 1] > 0
        f0_denorm = denorm_f0(f0, uv, pitch_cfg, pitch_padding=pitch_padding)
        f0[pitch_padding] = 0                            # in place, also on the caller's target (modules.py:934-935)
        pitch = f0_to_coarse(f0_denorm)
    emb = F.embedding(pitch, P[pre + "pitch_embed.weight"], padding_idx=0)
    return {"pitch_pred": pred, "f0_denorm": f0_denorm, "cwt": None, "f0_mean": None, "f0_std": None}, emb


def energy_embedding(P, cfg, x, target, control):
    """get_energy_embedding, modules.py:950-960."""
    pre = "variance_adaptor."
    pred = pitch_style_predictor(P, cfg, pre + "energy_predictor.", x).squeeze(-1)
    if target is not None:
        idx = torch.bucketize(target, P[pre + "energy_bins"])
    else:
        pred = pred * control
        idx = torch.bucketize(pred, P[pre + "energy_bins"])
    return pred, F.embedding(idx, P[pre + "energy_embedding.weight"], padding_idx=0)


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.GRU (batch_first), gates ordered r, z, n; runs over every (padded) step."""
    B, T, _ = x.shape
    Hd = w_hh.shape[1]
    gi = F.linear(x, w_ih, b_ih)
    h = x.new_zeros(B, Hd)
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = F.linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi[:, t].chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, 1), h


def parallel_prosody_predictor(P, cfg, pre, x, phoneme_level):
    """ParallelProsodyPredictor.forward, modules.py:630-648 (conv k3 + ReLU + LN(1e-5) twice, bi-GRU, bottleneck)."""
    k = cfg["prosody_modeling"]["liu2021"]["predictor_kernel_size"]
    E = x.shape[-1]
    h = x
    for i in (1, 2):
        pad = (k - 1) // 2 if i == 1 else 1
        h = F.conv1d(h.transpose(1, 2), P[pre + "conv_layer.conv1d_%d.conv.weight" % i],
                     P[pre + "conv_layer.conv1d_%d.conv.bias" % i], padding=pad).transpose(1, 2)
        h = F.layer_norm(F.relu(h), (E,), P[pre + "conv_layer.layer_norm_%d.weight" % i],
                         P[pre + "conv_layer.layer_norm_%d.bias" % i], 1e-5)
    fw, h_f = _gru_direction(h, P[pre + "gru.weight_ih_l0"], P[pre + "gru.weight_hh_l0"], P[pre + "gru.bias_ih_l0"],
                             P[pre + "gru.bias_hh_l0"], False)
    bw, h_b = _gru_direction(h, P[pre + "gru.weight_ih_l0_reverse"], P[pre + "gru.weight_hh_l0_reverse"],
                             P[pre + "gru.bias_ih_l0_reverse"], P[pre + "gru.bias_hh_l0_reverse"], True)
    vec = torch.cat([fw, bw], -1) if phoneme_level else torch.cat([h_f, h_b], -1).unsqueeze(1)
    return F.linear(vec, P[pre + "predictor_bottleneck.weight"], P[pre + "predictor_bottleneck.bias"])


def _batch_norm_train(h, P, pre, stats_out=None, momentum=0.1, eps=1e-5):
    """nn.BatchNorm{1,2}d in training mode: normalise with the batch mean / biased variance over every dim but the
    channel one; if `stats_out` is a dict it receives the buffers the module would hold afterwards
    (running = (1 - momentum) * running + momentum * batch statistic, the variance one unbiased; num_batches_tracked + 1)."""
    y = F.batch_norm(h, None, None, P[pre + "weight"], P[pre + "bias"], True, momentum, eps)
    if stats_out is not None:
        with torch.no_grad():
            dims = [d for d in range(h.dim()) if d != 1]
            n = h.numel() // h.shape[1]
            mean = h.mean(dims)
            var = h.var(dims, unbiased=False) * (n / max(n - 1, 1))
            stats_out[pre + "running_mean"] = (1 - momentum) * P[pre + "running_mean"] + momentum * mean
            stats_out[pre + "running_var"] = (1 - momentum) * P[pre + "running_var"] + momentum * var
            stats_out[pre + "num_batches_tracked"] = P[pre + "num_batches_tracked"] + 1
    return y


# --- liu2021 reference encoders: mel -> prosody, training mode only (modules.py:332-569) -----------
def _add_coords_2d(x):
    """AddCoords(rank=2, with_r=True), coordconv.py:36-71: channels [x, row coordinate, column coordinate, radius],
    coordinates scaled to [-1, 1], radius measured from (0.5, 0.5) as the reference does."""
    N, _, H, W = x.shape
    rows = torch.arange(H, dtype=torch.int32)[None, None, :, None].expand(1, 1, H, W)
    cols = torch.arange(W, dtype=torch.int32)[None, None, None, :].expand(1, 1, H, W)
    xx = (rows.float() / (H - 1)) * 2 - 1
    yy = (cols.float() / (W - 1)) * 2 - 1
    xx, yy = xx.repeat(N, 1, 1, 1), yy.repeat(N, 1, 1, 1)
    rr = torch.sqrt(torch.pow(xx - 0.5, 2) + torch.pow(yy - 0.5, 2))
    return torch.cat([x, xx, yy, rr], dim=1)


def reference_encoder(P, pre, pcfg, cfg, mel, mask, stats_out=None):
    """ReferenceEncoder.forward, modules.py:370-392: CoordConv2d + 5 Conv2d (stride (1, 2): time is kept), each followed
    by BatchNorm2d with BATCH statistics (the module only ever runs in training mode) and ReLU, then a GRU over time.
    Returns (memory [N, Ty, g], last hidden state [N, g])."""
    c = cfg["prosody_modeling"]["liu2021"]
    n_mel = pcfg["preprocessing"]["mel"]["n_mel_channels"]
    stride, pad = tuple(c["ref_enc_strides"]), tuple(c["ref_enc_pad"])
    N = mel.shape[0]
    out = mel.reshape(N, 1, -1, n_mel)
    for i in range(len(c["ref_enc_filters"])):
        if i == 0:
            out = F.conv2d(_add_coords_2d(out), P[pre + "convs.0.conv.weight"], P[pre + "convs.0.conv.bias"], stride, pad)
        else:
            out = F.conv2d(out, P[pre + "convs.%d.weight" % i], P[pre + "convs.%d.bias" % i], stride, pad)
        out = _batch_norm_train(out, P, pre + "bns.%d." % i, stats_out)
        out = F.relu(out)
    out = out.transpose(1, 2)
    out = out.contiguous().view(N, out.shape[1], -1)
    if mask is not None:
        out = out.masked_fill(mask.unsqueeze(-1), 0)
    return _gru_direction(out, P[pre + "gru.weight_ih_l0"], P[pre + "gru.weight_hh_l0"], P[pre + "gru.bias_ih_l0"],
                          P[pre + "gru.bias_hh_l0"], False)


def utterance_prosody_encoder(P, pcfg, cfg, mel, mel_mask, stats_out=None):
    """UtteranceLevelProsodyEncoder.forward, modules.py:555-569 (+ STL / StyleEmbedAttention with one head,
    modules.py:471-533; ref_attention_dropout is dropout: identity at p = 0)."""
    pre = "variance_adaptor.utterance_prosody_encoder."
    E = cfg["transformer"]["encoder_hidden"]
    _, h = reference_encoder(P, pre + "encoder.", pcfg, cfg, mel, mel_mask, stats_out)
    query = F.linear(h, P[pre + "encoder_prj.weight"], P[pre + "encoder_prj.bias"]).unsqueeze(1)       # [N, 1, E/2]
    tokens = torch.tanh(P[pre + "stl.embed"]).unsqueeze(0).expand(mel.shape[0], -1, -1)                # [N, tokens, E]
    values = F.linear(tokens, P[pre + "stl.attention.W_value.weight"])
    querys = F.linear(query, P[pre + "stl.attention.W_query.weight"])
    keys = F.linear(tokens, P[pre + "stl.attention.W_key.weight"])
    scores = F.softmax(torch.matmul(querys, keys.transpose(1, 2)) / (E ** 0.5), dim=2)
    style = torch.matmul(scores, values)                                                                # [N, 1, E]
    return F.linear(style, P[pre + "encoder_bottleneck.weight"], P[pre + "encoder_bottleneck.bias"])


def phoneme_prosody_encoder(P, pcfg, cfg, x, src_mask, mel, mel_mask, stats_out=None):
    """PhonemeLevelProsodyEncoder.forward, modules.py:421-450: text queries attend over the reference-encoder memory."""
    pre = "variance_adaptor.phoneme_prosody_encoder."
    E = cfg["transformer"]["encoder_hidden"]
    memory, _ = reference_encoder(P, pre + "encoder.", pcfg, cfg, mel, mel_mask, stats_out)
    emb = F.linear(memory, P[pre + "encoder_prj.weight"], P[pre + "encoder_prj.bias"])
    k, v = torch.split(emb, E, dim=-1)
    S, M = x.shape[1], mel.shape[1]
    q = F.linear(x, P[pre + "linears.0.linear.weight"])
    k = F.linear(k, P[pre + "linears.1.linear.weight"])
    attn = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(E)
    attn = attn.masked_fill(mel_mask.unsqueeze(1).expand(-1, S, -1), -float("inf"))
    attn = F.softmax(attn, dim=-1)
    attn = attn.masked_fill(src_mask.unsqueeze(-1).expand(-1, -1, M), 0.0)
    out = F.linear(torch.bmm(attn, v), P[pre + "encoder_bottleneck.weight"], P[pre + "encoder_bottleneck.bias"])
    return out.masked_fill(src_mask.unsqueeze(-1), 0.0), attn


# --- unsupervised duration modelling: AlignmentEncoder + MAS -----------------------------------
def alignment_encoder(P, queries, keys, src_mask, attn_prior, temperature, speaker_embed=None):
    """AlignmentEncoder.forward, modules.py:1176-1213.  queries [B,80,M] (mel), keys [B,C,S]."""
    pre = "variance_adaptor.aligner."
    if speaker_embed is not None:
        keys = keys + F.linear(speaker_embed, P[pre + "key_spk_proj.linear.weight"])[:, :, None]
        queries = queries + F.linear(speaker_embed, P[pre + "query_spk_proj.linear.weight"])[:, :, None]
    k = F.conv1d(keys, P[pre + "key_proj.0.conv.weight"], P[pre + "key_proj.0.conv.bias"], padding=1)
    k = F.conv1d(F.relu(k), P[pre + "key_proj.2.conv.weight"], P[pre + "key_proj.2.conv.bias"])
    q = F.conv1d(queries, P[pre + "query_proj.0.conv.weight"], P[pre + "query_proj.0.conv.bias"], padding=1)
    q = F.conv1d(F.relu(q), P[pre + "query_proj.2.conv.weight"], P[pre + "query_proj.2.conv.bias"])
    q = F.conv1d(F.relu(q), P[pre + "query_proj.4.conv.weight"], P[pre + "query_proj.4.conv.bias"])
    attn = (q[:, :, :, None] - k[:, :, None]) ** 2
    attn = -temperature * attn.sum(1, keepdim=True)
    if attn_prior is not None:
        attn = F.log_softmax(attn, dim=3) + torch.log(attn_prior[:, None] + 1e-8)
    logprob = attn.clone()
    attn = attn.masked_fill(src_mask[:, None, None, :], float("-inf"))
    return F.softmax(attn, dim=3), logprob


def mas_width1(attn_map):
    """mas_width1, modules.py:36-64 (numba in the reference; plain numpy loops here, small cases).
    attn_map: [M, S] probabilities.  Returns the 0/1 monotonic path matrix."""
    M, S = attn_map.shape
    with np.errstate(divide="ignore"):
        a = np.log(attn_map)
    a[0, 1:] = -np.inf
    log_p = np.zeros_like(a)
    log_p[0] = a[0]
    prev = np.zeros((M, S), dtype=np.int64)
    for i in range(1, M):
        left = np.concatenate([[-np.inf], log_p[i - 1, :-1]]).astype(a.dtype)
        take_left = left >= log_p[i - 1]
        take_left[0] = False
        best = np.where(take_left, left, log_p[i - 1])
        log_p[i] = a[i] + best
        prev[i] = np.arange(S) - take_left.astype(np.int64)
    opt = np.zeros_like(a)
    j = S - 1
    for i in range(M - 1, -1, -1):
        opt[i, j] = 1
        j = prev[i, j]
    opt[0, j] = 1
    return opt


def binarize_attention(attn, in_lens, out_lens):
    """binarize_attention_parallel / b_mas, modules.py:66-75, 863-872."""
    a = attn.detach().cpu().numpy()
    out = np.zeros_like(a)
    for b in range(a.shape[0]):
        m, s = int(out_lens[b]), int(in_lens[b])
        out[b, 0, :m, :s] = mas_width1(a[b, 0, :m, :s].copy())
    return torch.from_numpy(out)


def phoneme_level_energy(duration, src_len, energy_frame):
    """VarianceAdaptor.get_phoneme_level_energy (modules.py:882-888) + utils/tools.py:56-66 + pad_1D."""
    B = duration.shape[0]
    rows = []
    for b in range(B):
        d = duration[b, : int(src_len[b])].int().numpy()
        e = energy_frame[b].numpy().copy()
        pos = 0
        for i, di in enumerate(d):
            e[i] = np.mean(e[pos: pos + di]) if di > 0 else 0
            pos += di
        rows.append(e[: len(d)])
    L = max(len(r) for r in rows)
    out = np.zeros((B, L), dtype=np.float32)
    for b, r in enumerate(rows):
        out[b, : len(r)] = r
    return torch.from_numpy(out)


def variance_adaptor(P, pcfg, cfg, tcfg, speaker_embedding, text, text_embedding, src_len, src_mask, mel, mel_len,
                     mel_mask, max_len, pitch_target, energy_target, duration_target, attn_prior,
                     p_control, e_control, d_control, step, training=False, stats_out=None):
    """VarianceAdaptor.forward (prosody model 'none' or 'liu2021'), modules.py:962-1114.  `training` selects the
    reference encoders of liu2021 (modules.py:1005-1016); every dropout is the identity (parity is defined at p = 0).
    The predictor inputs carry the reference's gradient scaling x.detach() + predictor_grad * (x - x.detach())
    (values unchanged), so autograd through this function reproduces the reference's gradients."""
    assert cfg["prosody_modeling"]["model_type"] in ("none", "liu2021")
    pitch_type = pcfg["preprocessing"]["pitch"]["pitch_type"]
    assert pitch_type in ("cwt", "frame", "ph") and not pcfg["preprocessing"]["pitch"].get("pitch_ar", False)
    learn_alignment = cfg["duration_modeling"]["learn_alignment"]
    x = text.clone()
    if speaker_embedding is not None:
        x = x + speaker_embedding.unsqueeze(1)
    prosody_info = None
    if cfg["prosody_modeling"]["model_type"] == "liu2021":
        # eval mode: the predictors stand in for the reference encoders (modules.py:1002-1023)
        u_emb = p_emb_ref = p_attn = None
        if training:
            u_emb = utterance_prosody_encoder(P, pcfg, cfg, mel, mel_mask, stats_out)
            p_emb_ref, p_attn = phoneme_prosody_encoder(P, pcfg, cfg, x, src_mask, mel, mel_mask, stats_out)
        u_vec = parallel_prosody_predictor(P, cfg, "variance_adaptor.utterance_prosody_predictor.", x, False)
        x = x + F.linear(u_emb if training else u_vec, P["variance_adaptor.utterance_prosody_prj.weight"],
                         P["variance_adaptor.utterance_prosody_prj.bias"])
        p_vec = parallel_prosody_predictor(P, cfg, "variance_adaptor.phoneme_prosody_predictor.", x, True)
        x = x + F.linear(p_emb_ref if training else p_vec, P["variance_adaptor.phoneme_prosody_prj.weight"],
                         P["variance_adaptor.phoneme_prosody_prj.bias"])
        prosody_info = (u_emb, p_emb_ref, u_vec, p_vec, p_attn)
    g = cfg["variance_predictor"]["predictor_grad"]
    log_d = duration_predictor(P, cfg, x.detach() + g * (x - x.detach()), src_mask)

    attn_soft = attn_hard = attn_hard_dur = attn_logprob = None
    if attn_prior is not None:
        assert learn_alignment and duration_target is None and mel is not None
        attn_soft, attn_logprob = alignment_encoder(
            P, mel.transpose(1, 2), text_embedding.transpose(1, 2), src_mask, attn_prior.transpose(1, 2),
            cfg["duration_modeling"]["aligner_temperature"], speaker_embedding)
        attn_hard = binarize_attention(attn_soft, src_len, mel_len)
        attn_hard_dur = attn_hard.sum(2)[:, 0, :]
    attn_out = (attn_soft, attn_hard, attn_hard_dur, attn_logprob)

    x_org = x.clone()
    if attn_prior is not None:
        if step < tcfg["duration"]["binarization_start_steps"]:
            x = torch.bmm(attn_soft.squeeze(1), x)
        else:
            x, mel_len = length_regulate(x, attn_hard_dur, max_len)
        duration_rounded = attn_hard_dur
        pitch_target["mel2ph"] = durations_to_mel2ph(duration_rounded, src_mask)[:, :max_len]
    elif duration_target is not None:
        assert not learn_alignment
        x, mel_len = length_regulate(x, duration_target, max_len)
        duration_rounded = duration_target
    else:
        duration_rounded = torch.clamp(torch.round(torch.exp(log_d) - 1) * d_control, min=0)
        x, mel_len = length_regulate(x, duration_rounded, max_len)
        mel_mask = pad_mask_from_lengths(mel_len)
        mel2ph = durations_to_mel2ph(duration_rounded, src_mask)

    x_sum = x.clone()
    pitch_pred = energy_pred = None
    if cfg["variance_embedding"]["use_pitch_embed"]:
        embed = pitch_embedding_cwt if pitch_type == "cwt" else pitch_embedding_frame_or_ph
        if pitch_target is not None:
            mel2ph = pitch_target["mel2ph"]
            if pitch_type == "cwt":
                pitch_target["f0"] = cwt_to_f0_norm(pitch_target["cwt_spec"], pitch_target["f0_mean"],
                                                    pitch_target["f0_std"], mel2ph.shape[1],
                                                    pcfg["preprocessing"]["pitch"])
                pitch_target["f0_cwt"] = pitch_target["f0"]
            if pitch_type == "ph":
                pitch_target["f0"] = phoneme_level_pitch(text, src_len, mel2ph, mel_len, pitch_target["f0"])
            pitch_pred, p_emb = embed(P, cfg, pcfg, x, pitch_target["f0"], pitch_target["uv"], mel2ph, p_control, x_org)
        else:
            pitch_pred, p_emb = embed(P, cfg, pcfg, x, None, None, mel2ph, p_control, x_org)
        x_sum = x_sum + p_emb
    if cfg["variance_embedding"]["use_energy_embed"]:
        level = pcfg["preprocessing"]["energy"]["feature"]
        if level == "frame_level":
            energy_pred, e_emb = energy_embedding(P, cfg, x, energy_target, e_control)
            x_sum = x_sum + e_emb
        else:
            if attn_prior is not None:
                energy_target = phoneme_level_energy(attn_hard_dur, src_len, energy_target)
            energy_pred, e_emb = energy_embedding(P, cfg, x_org, energy_target, e_control)
            x_sum = x_sum + length_regulate(e_emb, duration_rounded, max_len)[0]
    return (x_sum, pitch_target, pitch_pred, energy_target, energy_pred, log_d, duration_rounded, mel_len, mel_mask,
            attn_out, prosody_info)


# ----------------------------------------------------------------------------------------------
# mel head (model/CompTransTTS.py:133-135, model/modules.py:78-148)
# ----------------------------------------------------------------------------------------------
def postnet(P, x, training=False, stats_out=None):
    """PostNet.forward, modules.py:140-148.  Eval: BatchNorm1d running statistics.  Training: batch statistics over all
    B x T frames, padded ones included, as the reference computes them (`stats_out` receives the updated buffers); the
    hard-coded dropout(0.5) is the identity (parity is defined at p = 0)."""
    h = x.transpose(1, 2)
    for i in range(5):
        pre = "postnet.convolutions.%d." % i
        h = F.conv1d(h, P[pre + "0.conv.weight"], P[pre + "0.conv.bias"], padding=2)
        if training:
            h = _batch_norm_train(h, P, pre + "1.", stats_out)
        else:
            h = F.batch_norm(h, P[pre + "1.running_mean"], P[pre + "1.running_var"], P[pre + "1.weight"],
                             P[pre + "1.bias"], False, 0.1, 1e-5)
        if i < 4:
            h = torch.tanh(h)
    return h.transpose(1, 2)


# ----------------------------------------------------------------------------------------------
# top level (model/CompTransTTS.py:64-152)
# ----------------------------------------------------------------------------------------------
def comp_trans_tts_forward(P, pcfg, cfg, tcfg, speakers, texts, src_lens, max_src_len, mels=None, mel_lens=None,
                           max_mel_len=None, p_targets=None, e_targets=None, d_targets=None, attn_priors=None,
                           spker_embeds=None, p_control=1.0, e_control=1.0, d_control=1.0, step=None, taps=None,
                           training=False, stats_out=None):
    """Returns the reference's 14-tuple.  `taps`, if a dict, receives intermediate activations.
    training=True restates model.train() with every dropout probability 0: PostNet / conformer / reference-encoder
    BatchNorm on batch statistics (`stats_out`, a dict, receives the updated running buffers), liu2021 reference
    encoders; differentiable, so torch.autograd through it is the oracle for the backward pass."""
    block = cfg["block_type"]
    src_masks = pad_mask_from_lengths(src_lens, max_src_len)
    mel_masks = pad_mask_from_lengths(mel_lens, max_mel_len) if mel_lens is not None else None
    if block == "transformer_fs2":
        enc, word = encoder_fs2(P, cfg, texts, src_masks, taps)
    else:
        from . import ctts_oracle_blocks as OB
        enc, word = OB.ENCODERS[block](P, cfg, texts, src_masks, taps, training=training, stats_out=stats_out)
    if taps is not None:
        taps["encoder_out"] = enc
    spk = None
    if cfg["multi_speaker"]:
        if pcfg["preprocessing"]["speaker_embedder"] == "none":
            spk = F.embedding(speakers, P["speaker_emb.weight"])
        else:
            assert spker_embeds is not None, "Speaker embedding should not be None"
            spk = F.linear(spker_embeds, P["speaker_emb.weight"], P["speaker_emb.bias"])
    (x, p_targets, p_pred, e_targets, e_pred, log_d, d_rounded, mel_lens, mel_masks, attn_outs, prosody) = \
        variance_adaptor(P, pcfg, cfg, tcfg, spk, enc, word, src_lens, src_masks, mels, mel_lens, mel_masks,
                         max_mel_len, p_targets, e_targets, d_targets, attn_priors, p_control, e_control, d_control,
                         step, training, stats_out)
    if taps is not None:
        taps["decoder_in"] = x
    if block == "transformer_fs2":
        dec, mel_masks = decoder_fs2(P, cfg, x, mel_masks, taps)
    else:
        from . import ctts_oracle_blocks as OB
        dec, mel_masks = OB.DECODERS[block](P, cfg, x, mel_masks, taps, training=training, stats_out=stats_out)
    if taps is not None:
        taps["decoder_out"] = dec
    mel = F.linear(dec, P["mel_linear.weight"], P["mel_linear.bias"])
    post = postnet(P, mel, training, stats_out) + mel
    return (mel, post, p_pred, e_pred, log_d, d_rounded, src_masks, mel_masks, src_lens, mel_lens, attn_outs, prosody,
            p_targets, e_targets)
