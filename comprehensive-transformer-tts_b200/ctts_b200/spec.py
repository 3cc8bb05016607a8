"""Parameter inventory of the acoustic model: state_dict names, shapes and initialisers.

The drop-in contract is that a checkpoint written by the reference's
`train.py:190-200` loads with `load_state_dict(strict=True)` and vice versa, so every tensor
name and shape the reference registers has to exist here.  Instead of re-creating the
reference's nn.Module class tree, the names are generated from this table and hung on anonymous
container modules (`module.py`); the forward pass never walks the tree, it reads a flat
`name -> tensor` dict.

Entry = (name, shape, kind, init):
  kind  "param" (trainable) | "frozen" (Parameter, requires_grad=False) | "buffer"
  init  "xavier[:gain]" | "default:<fan_in>" (torch's Conv/Linear default: U(+-1/sqrt(fan_in)))
        | "emb:<dim>" (N(0, dim^-0.5), row 0 zero) | "emb1" (N(0,1), row 0 zero: nn.Embedding(padding_idx=0))
        | "ones" | "zeros" | "sinusoid_interleaved" | "linspace:a:b" | "normal01" | "normal002" | "count"
        | "tie:<name>" (the same Parameter object as <name>)
References: model/transformers/transformer_fs2.py:16-45,75-134,154-218,278-330;
model/modules.py:78-138,726-861,1117-1174,1252-1298,1313-1341; model/CompTransTTS.py:34-62.
"""
import json
import math
import os

VOCAB_DEFAULT = 361  # len(text.symbols.symbols) + 1, transformer_fs2.py:91-93 (probed: 360 symbols)


def vocab_size():
    try:  # inside the reference tree the real symbol table wins
        from text.symbols import symbols  # type: ignore
        return len(symbols) + 1
    except Exception:
        return VOCAB_DEFAULT


def _lin(out, name, o, i, bias=True, init=None):
    out.append((name + ".weight", (o, i), "param", init or "default:%d" % i))
    if bias:
        out.append((name + ".bias", (o,), "param", "zeros" if init else "default:%d" % i))


def _conv(out, name, o, i, k, init=None):
    out.append((name + ".weight", (o, i, k), "param", init or "default:%d" % (i * k)))
    out.append((name + ".bias", (o,), "param", "default:%d" % (i * k)))


def _ln(out, name, c):
    out.append((name + ".weight", (c,), "param", "ones"))
    out.append((name + ".bias", (c,), "param", "zeros"))


def _bn(out, name, c):
    _ln(out, name, c)
    out.append((name + ".running_mean", (c,), "buffer", "zeros"))
    out.append((name + ".running_var", (c,), "buffer", "ones"))
    out.append((name + ".num_batches_tracked", (), "buffer", "count"))


# ---------------------------------------------------------------------------------------------
def _fs2_stack(out, pre, n_layers, c, k):
    f = 4 * c
    for i in range(n_layers):
        p = "%slayers.%d.op." % (pre, i)
        _ln(out, p + "layer_norm1", c)
        out.append((p + "self_attn.in_proj_weight", (3 * c, c), "param", "xavier"))
        out.append((p + "self_attn.out_proj.weight", (c, c), "param", "xavier"))
        _ln(out, p + "layer_norm2", c)
        _conv(out, p + "ffn.ffn_1", f, c, k)
        _lin(out, p + "ffn.ffn_2", c, f, init="xavier")
    _ln(out, pre + "layer_norm", c)


def _blocks_fs2(out, cfg):
    c = cfg["transformer_fs2"]
    d = c["encoder_hidden"]
    _fs2_stack(out, "encoder.", c["encoder_layer"], d, c["ffn_kernel_size"])
    out.append(("encoder.embed_tokens.weight", (vocab_size(), d), "param", "emb:%d" % d))
    out.append(("encoder.embed_positions._float_tensor", (1,), "buffer", "zeros"))
    d2 = c["decoder_hidden"]
    out.append(("decoder.pos_embed_alpha", (1,), "param", "ones"))
    out.append(("decoder.embed_positions._float_tensor", (1,), "buffer", "zeros"))
    _fs2_stack(out, "decoder.", c["decoder_layer"], d2, c["ffn_kernel_size"])
    return d, d2


def _predictor(out, pre, idim, chans, layers, k, odim, with_pos):
    if with_pos:
        out.append((pre + "pos_embed_alpha", (1,), "param", "ones"))
    for l in range(layers):
        _conv(out, "%sconv.%d.1" % (pre, l), chans, idim if l == 0 else chans, k)
        _ln(out, "%sconv.%d.3" % (pre, l), chans)
    _lin(out, pre + "linear", odim, chans)
    if with_pos:
        out.append((pre + "embed_positions._float_tensor", (1,), "buffer", "zeros"))


def energy_range(preprocess_config, model_config):
    """(min, max) of the energy statistics the bins are built from, modules.py:787-799."""
    learn = model_config["duration_modeling"]["learn_alignment"]
    level = preprocess_config["preprocessing"]["energy"]["feature"]
    assert level in ("frame_level", "phoneme_level")
    tag = "phone" if (not learn and level == "phoneme_level") else "frame"  # utils/tools.py:30-44
    with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "stats.json")) as f:
        stats = json.load(f)
    lo, hi = stats["energy_%s_%s" % ("unsup" if learn else "sup", tag)][:2]
    return float(lo), float(hi)


def _variance_adaptor(out, pcfg, cfg, d_model):
    vp, ve = cfg["variance_predictor"], cfg["variance_embedding"]
    hid = cfg["transformer"]["encoder_hidden"]  # modules.py:739 reads the `transformer` section
    pre = "variance_adaptor."
    pitch = pcfg["preprocessing"]["pitch"]
    if ve["use_energy_embed"]:
        lo, hi = energy_range(pcfg, cfg)
        if ve["energy_quantization"] == "log":
            init = "logspace:%r:%r" % (lo, hi)
        else:
            init = "linspace:%r:%r" % (lo, hi)
        out.append((pre + "energy_bins", (ve["energy_n_bins"] - 1,), "frozen", init))
    _predictor(out, pre + "duration_predictor.", hid, vp["filter_size"], vp["dur_predictor_layers"],
               vp["dur_predictor_kernel"], 1, False)
    if ve["use_pitch_embed"]:
        if pitch["pitch_type"] not in ("cwt", "frame", "ph"):
            raise NotImplementedError("pitch_type %r" % pitch["pitch_type"])
        if pitch["pitch_type"] == "cwt":
            h = vp["cwt_hidden_size"]
            _lin(out, pre + "cwt_predictor.0", h, hid)
            _predictor(out, pre + "cwt_predictor.1.", h, vp["filter_size"], vp["predictor_layers"], vp["predictor_kernel"],
                       11 if pitch["use_uv"] else 10, True)
            _lin(out, pre + "cwt_stats_layers.0", h, hid)
            _lin(out, pre + "cwt_stats_layers.2", h, h)
            _lin(out, pre + "cwt_stats_layers.4", 2, h)
        else:   # modules.py:778-785: one PitchPredictor on the frame- ('frame': f0 + uv) or phoneme-level ('ph': f0) input
            if pitch.get("pitch_ar"):
                raise NotImplementedError("pitch_ar (autoregressive pitch predictor, modules.py:923-926) is not built")
            _predictor(out, pre + "pitch_predictor.", hid, vp["filter_size"], vp["predictor_layers"], vp["predictor_kernel"],
                       2 if pitch["pitch_type"] == "frame" else 1, True)
        out.append((pre + "pitch_embed.weight", (ve["pitch_n_bins"], hid), "param", "emb:%d" % hid))
    if ve["use_energy_embed"]:
        _predictor(out, pre + "energy_predictor.", hid, vp["filter_size"], vp["predictor_layers"],
                   vp["predictor_kernel"], 1, True)
        out.append((pre + "energy_embedding.weight", (ve["energy_n_bins"], hid), "param", "emb:%d" % hid))
    if cfg["duration_modeling"]["learn_alignment"]:
        mel = pcfg["preprocessing"]["mel"]["n_mel_channels"]
        a = pre + "aligner."
        _conv(out, a + "key_proj.0.conv", 2 * d_model, d_model, 3, init="xavier:relu")
        _conv(out, a + "key_proj.2.conv", mel, 2 * d_model, 1, init="xavier")
        _conv(out, a + "query_proj.0.conv", 2 * mel, mel, 3, init="xavier:relu")
        _conv(out, a + "query_proj.2.conv", mel, 2 * mel, 1, init="xavier")
        _conv(out, a + "query_proj.4.conv", mel, mel, 1, init="xavier")
        if cfg["multi_speaker"]:
            out.append((a + "key_spk_proj.linear.weight", (d_model, d_model), "param", "xavier"))
            out.append((a + "query_spk_proj.linear.weight", (mel, d_model), "param", "xavier"))
    model_type = cfg["prosody_modeling"]["model_type"]
    if model_type == "liu2021":
        _liu2021(out, pcfg, cfg)
    elif model_type != "none":
        # du2021 (GMM-MDN) is out of scope: SURVEY.md section 2 row 2
        raise NotImplementedError("prosody_modeling.model_type %r is not built" % model_type)


def _liu2021(out, pcfg, cfg):
    """Implicit prosody modelling (Liu et al. 2021): modules.py:332-648, 840-861; coordconv.py:140-159.
    The two reference encoders (mel -> prosody) only run in training mode (modules.py:1005-1007) but their
    parameters are part of every checkpoint, so they are registered here as well."""
    c = cfg["prosody_modeling"]["liu2021"]
    E = cfg["transformer"]["encoder_hidden"]
    mel = pcfg["preprocessing"]["mel"]["n_mel_channels"]
    g = c["ref_enc_gru_size"]
    filt = [1] + list(c["ref_enc_filters"])
    kh, kw = c["ref_enc_size"]
    L = mel
    for _ in range(len(filt) - 1):      # ReferenceEncoder.calculate_channels(L, 3, 2, 1, K)
        L = (L - 3 + 2 * 1) // 2 + 1
    for enc in ("utterance_prosody_encoder", "phoneme_prosody_encoder"):
        p = "variance_adaptor.%s.encoder." % enc
        for i in range(len(filt) - 1):
            fan = filt[i] * kh * kw
            out.append((p + "convs.%d.weight" % i, (filt[i + 1], filt[i], kh, kw), "param", "default:%d" % fan))
            out.append((p + "convs.%d.bias" % i, (filt[i + 1],), "param", "default:%d" % fan))
            if i == 0:  # CoordConv2d keeps the parent Conv2d's tensors AND its own conv (+2 coords +1 radius)
                fan = (filt[0] + 3) * kh * kw
                out.append((p + "convs.0.conv.weight", (filt[1], filt[0] + 3, kh, kw), "param", "default:%d" % fan))
                out.append((p + "convs.0.conv.bias", (filt[1],), "param", "default:%d" % fan))
        for i in range(len(filt) - 1):
            _bn(out, p + "bns.%d" % i, filt[i + 1])
        _gru(out, p + "gru", filt[-1] * L, g, False)
        q = "variance_adaptor.%s." % enc
        if enc == "utterance_prosody_encoder":
            _lin(out, q + "encoder_prj", E // 2, g)
            out.append((q + "stl.embed", (c["token_num"], E), "param", "normal05"))
            out.append((q + "stl.attention.W_query.weight", (E, E // 2), "param", "default:%d" % (E // 2)))
            out.append((q + "stl.attention.W_key.weight", (E, E), "param", "default:%d" % E))
            out.append((q + "stl.attention.W_value.weight", (E, E), "param", "default:%d" % E))
            _lin(out, q + "encoder_bottleneck", c["bottleneck_size_u"], E)
        else:
            out.append((q + "linears.0.linear.weight", (E, E), "param", "xavier"))
            out.append((q + "linears.1.linear.weight", (E, E), "param", "xavier"))
            _lin(out, q + "encoder_prj", 2 * E, g)
            _lin(out, q + "encoder_bottleneck", c["bottleneck_size_p"], E)
    for name, bott in (("utterance_prosody_predictor", c["bottleneck_size_u"]),
                       ("phoneme_prosody_predictor", c["bottleneck_size_p"])):
        p = "variance_adaptor.%s." % name
        k = c["predictor_kernel_size"]
        _conv(out, p + "conv_layer.conv1d_1.conv", E, E, k, init="xavier")
        _ln(out, p + "conv_layer.layer_norm_1", E)
        _conv(out, p + "conv_layer.conv1d_2.conv", E, E, k, init="xavier")
        _ln(out, p + "conv_layer.layer_norm_2", E)
        _gru(out, p + "gru", E, E // 2, True)
        _lin(out, p + "predictor_bottleneck", bott, E)
    _lin(out, "variance_adaptor.utterance_prosody_prj", E, c["bottleneck_size_u"])
    _lin(out, "variance_adaptor.phoneme_prosody_prj", E, c["bottleneck_size_p"])


def _gru(out, name, in_size, hidden, bidirectional):
    for suffix in ([""] + (["_reverse"] if bidirectional else [])):
        out.append((name + ".weight_ih_l0" + suffix, (3 * hidden, in_size), "param", "default:%d" % hidden))
        out.append((name + ".weight_hh_l0" + suffix, (3 * hidden, hidden), "param", "default:%d" % hidden))
        out.append((name + ".bias_ih_l0" + suffix, (3 * hidden,), "param", "default:%d" % hidden))
        out.append((name + ".bias_hh_l0" + suffix, (3 * hidden,), "param", "default:%d" % hidden))


def _postnet(out):
    chans = [80, 512, 512, 512, 512, 80]  # PostNet() is built with defaults, CompTransTTS.py:41
    for i in range(5):
        _conv(out, "postnet.convolutions.%d.0.conv" % i, chans[i + 1], chans[i], 5,
              init="xavier:tanh" if i < 4 else "xavier")
        _bn(out, "postnet.convolutions.%d.1" % i, chans[i + 1])


def parameter_spec(preprocess_config, model_config):
    """Ordered list of (name, shape, kind, init) for the configured model."""
    block = model_config["block_type"]
    out = []
    if block == "transformer_fs2":
        d_enc, d_dec = _blocks_fs2(out, model_config)
    elif block in ("transformer", "fastformer", "conformer"):
        from . import spec_blocks
        d_enc, d_dec = spec_blocks.BLOCK_SPECS[block](out, model_config)
    else:
        # lstransformer / reformer are out of scope
        raise NotImplementedError(block)
    _variance_adaptor(out, preprocess_config, model_config, d_enc)
    out.append(("mel_linear.weight", (preprocess_config["preprocessing"]["mel"]["n_mel_channels"], d_dec), "param",
                "default:%d" % d_dec))
    out.append(("mel_linear.bias", (preprocess_config["preprocessing"]["mel"]["n_mel_channels"],), "param",
                "default:%d" % d_dec))
    _postnet(out)
    if model_config["multi_speaker"]:
        kind = preprocess_config["preprocessing"]["speaker_embedder"]
        if kind == "none":
            with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "speakers.json")) as f:
                n_spk = len(json.load(f))
            out.append(("speaker_emb.weight", (n_spk, d_enc), "param", "normal01"))
        else:
            _lin(out, "speaker_emb", d_enc, model_config["external_speaker_dim"])
    names = [n for n, _, _, _ in out]
    assert len(set(names)) == len(names), "duplicate parameter names"
    return out, d_enc, d_dec


def gain_of(tag):
    return {"": 1.0, "relu": math.sqrt(2.0), "tanh": 5.0 / 3.0}[tag]
