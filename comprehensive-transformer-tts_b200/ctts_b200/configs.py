"""Built-in copies of the configuration dictionaries the model constructor consumes.

The reference loads three YAML files per dataset with `get_configs_of` (utils/tools.py:19-27,
config/{LJSpeech,VCTK}/*.yaml) and hands the resulting dicts to
`CompTransTTS(preprocess_config, model_config, train_config)`.  Dicts loaded from the
reference's own YAML files work unchanged with this package (that is the drop-in contract);
this module only exists so that tests / bench / smoke can build the same dicts on a machine where
the reference tree is absent.  Only keys read on the acoustic-model path are included
(SURVEY.md section 5 "Config").
"""
import copy
import os

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

_BLOCK = dict(encoder_layer=4, encoder_head=2, encoder_hidden=256, decoder_layer=6, decoder_head=2,
              decoder_hidden=256)


def _model(dataset):
    multi = dataset == "VCTK"
    cfg = {
        "block_type": "transformer_fs2",
        "duration_modeling": {"learn_alignment": True, "aligner_temperature": 0.0005},
        "prosody_modeling": {
            "model_type": "none",
            "liu2021": dict(bottleneck_size_u=256, bottleneck_size_p=4, ref_enc_filters=[32, 32, 64, 64, 128, 128],
                            ref_enc_size=[3, 3], ref_enc_strides=[1, 2], ref_enc_pad=[1, 1], ref_enc_gru_size=32,
                            ref_attention_dropout=0.0, token_num=32, predictor_kernel_size=3,
                            predictor_dropout=0.5),
        },
        "transformer_fs2": dict(_BLOCK, ffn_kernel_size=9, encoder_dropout=0.1, decoder_dropout=0.1),
        "transformer": dict(_BLOCK, conv_filter_size=1024, conv_kernel_size=[9, 1], encoder_dropout=0.2,
                            decoder_dropout=0.2),
        "conformer": dict(_BLOCK, encoder_head=8, decoder_head=8, feed_forward_expansion_factor=4,
                          conv_expansion_factor=2, conv_kernel_size=31, half_step_residual=True,
                          encoder_dropout=0.1, decoder_dropout=0.1),
        "variance_predictor": dict(filter_size=256, predictor_grad=0.1, predictor_layers=2, predictor_kernel=5,
                                   cwt_hidden_size=128, cwt_std_scale=0.8, dur_predictor_layers=2,
                                   dur_predictor_kernel=3, dropout=0.5, ffn_padding="SAME", ffn_act="gelu"),
        "variance_embedding": dict(use_pitch_embed=True, pitch_n_bins=300, use_energy_embed=True,
                                   energy_n_bins=256, energy_quantization="linear"),
        "multi_speaker": multi,
        "max_seq_len": 1500 if multi else 1000,
    }
    if multi:
        cfg["external_speaker_dim"] = 512
    return cfg


def _preprocess(dataset):
    pre = {
        "dataset": dataset,
        "path": {"preprocessed_path": os.path.join(_ASSETS, dataset)},
        "preprocessing": {
            "mel": {"n_mel_channels": 80},
            "pitch": dict(pitch_type="cwt", pitch_norm="log", pitch_norm_eps=1e-9, pitch_ar=False, with_f0=True,
                          with_f0cwt=True, use_uv=True, cwt_scales=0.01 * 2.0 ** np.arange(10)),
            "energy": {"feature": "phoneme_level", "normalization": True},
            "duration": {"beta_binomial_scaling_factor": 1.0},
        },
    }
    if dataset == "VCTK":
        pre["preprocessing"]["speaker_embedder"] = "DeepSpeaker"
    return pre


def _train(dataset):
    return {
        "seed": 1234,
        "loss": dict(noise_loss="l1", dur_loss="mse", pitch_loss="l1", cwt_loss="l1", lambda_f0=1.0, lambda_uv=1.0,
                     lambda_ph_dur=1.0, lambda_word_dur=0.0 if dataset == "VCTK" else 1.0, lambda_sent_dur=1.0),
        "step": dict(var_start_steps=50000),
        "duration": dict(binarization_start_steps=6000, binarization_loss_enable_steps=18000,
                         binarization_loss_warmup_steps=10000),
        "prosody": dict(gmm_mdn_beta=0.02, prosody_loss_enable_steps=100000),
    }


def builtin_configs(dataset="LJSpeech", block_type=None, learn_alignment=None, prosody=None, pitch_type=None):
    """(preprocess_config, model_config, train_config) equal, on every key this path reads, to the
    reference's config/<dataset>/*.yaml (after train.py:229-231 patched `cwt_scales` in)."""
    if dataset not in ("LJSpeech", "VCTK"):
        raise ValueError("unknown dataset %r" % (dataset,))
    p, m, t = _preprocess(dataset), _model(dataset), _train(dataset)
    if block_type is not None:
        m["block_type"] = block_type
    if learn_alignment is not None:
        m["duration_modeling"]["learn_alignment"] = bool(learn_alignment)
    if prosody is not None:
        m["prosody_modeling"]["model_type"] = prosody
    if pitch_type is not None:
        p["preprocessing"]["pitch"]["pitch_type"] = pitch_type
    return copy.deepcopy(p), copy.deepcopy(m), copy.deepcopy(t)
