"""Definition of the golden cases (shared by make_golden.py, the CPU oracle tests and the GPU parity tests).

Every case is fully determined by this table: weights come from `synth.synthetic_state_dict`
(seeded by tensor name), inputs from `synth.ljspeech_batch` (seeded), so the fixtures only have to
store the REFERENCE's outputs.
"""
CASES = {
    # BASELINE.json configs[0]: transformer_fs2, LJSpeech config, batch 2, seq_len ~100, free-running inference.
    # Durations are pinned to 3 frames / phoneme (the random-init duration predictor would emit ~0 frames).
    "fs2_infer_c1": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="infer",
                         batch=2, s_max=100, s_step=9, pin=3, seed=0),
    # supervised / teacher-forced branch (modules.py:1054-1057): explicit durations U[1,15], pitch / energy targets
    "fs2_teacher": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                        batch=3, s_max=40, s_step=7, pin=None, seed=1),
    # post-LN "transformer" block (model/transformers/transformer.py): free-running and teacher-forced
    "transformer_infer": dict(dataset="LJSpeech", block_type="transformer", learn_alignment=False, mode="infer",
                              batch=2, s_max=60, s_step=11, pin=4, seed=2),
    "transformer_teacher": dict(dataset="LJSpeech", block_type="transformer", learn_alignment=False, mode="teacher",
                                batch=2, s_max=30, s_step=7, pin=None, seed=3),
    # fastformer (additive attention with the reference's inverted mask; padded AND unpadded utterances in one batch)
    "fastformer_infer": dict(dataset="LJSpeech", block_type="fastformer", learn_alignment=False, mode="infer",
                             batch=2, s_max=60, s_step=11, pin=4, seed=4),
    # conformer (relative-position attention without padding mask, GLU + depthwise conv + BatchNorm + Swish)
    "conformer_infer": dict(dataset="LJSpeech", block_type="conformer", learn_alignment=False, mode="infer",
                            batch=2, s_max=60, s_step=11, pin=4, seed=5),
    "conformer_teacher": dict(dataset="LJSpeech", block_type="conformer", learn_alignment=False, mode="teacher",
                              batch=2, s_max=30, s_step=7, pin=None, seed=6),
    # unsupervised duration modelling: AlignmentEncoder + MAS + hard-duration upsampling + phoneme-level energy from
    # frame-level targets (learn_alignment True, attn_priors given, step 120000 > binarization_start_steps)
    "fs2_unsup": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=True, mode="unsup",
                      batch=3, s_max=30, s_step=7, pin=None, seed=7),
    # BASELINE configs[3] family: fastformer, VCTK config (multi-speaker, DeepSpeaker 512-d embeddings -> Linear),
    # unsupervised alignment with speaker-conditioned aligner projections
    "fastformer_vctk_unsup": dict(dataset="VCTK", block_type="fastformer", learn_alignment=True, mode="unsup",
                                  batch=3, s_max=24, s_step=5, pin=None, seed=8),
    # BASELINE configs[4] family: transformer_fs2 + liu2021 implicit prosody (eval: conv + bi-GRU predictors)
    "fs2_liu2021_infer": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="infer",
                              batch=2, s_max=40, s_step=9, pin=4, seed=9, prosody="liu2021"),
    # the non-default pitch parametrisations (preprocess.yaml pitch_type): one PitchPredictor on the frame-level input
    # ('frame': f0 + voiced/unvoiced logit per frame) or on the phoneme-level input ('ph': one f0 per phoneme, gathered to
    # frames through mel2ph), modules.py:890-906,927-938
    "fs2_pitch_frame_infer": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="infer",
                                  batch=2, s_max=40, s_step=9, pin=4, seed=20, pitch_type="frame"),
    "fs2_pitch_frame_teacher": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                                    batch=2, s_max=30, s_step=7, pin=None, seed=21, pitch_type="frame"),
    "fs2_pitch_ph_infer": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="infer",
                               batch=2, s_max=40, s_step=9, pin=4, seed=22, pitch_type="ph"),
    "fs2_pitch_ph_teacher": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                                 batch=2, s_max=30, s_step=7, pin=None, seed=23, pitch_type="ph"),
}

# Training-step cases (make_golden_train.py / test_oracle_train.py): model.train() with every dropout probability 0,
# teacher-forced targets, a fixed scalar objective (train_objective below) and its gradients w.r.t. every parameter.
TRAIN_CASES = {
    "fs2_train": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                      batch=3, s_max=24, s_step=5, pin=None, seed=11),
    # liu2021: the two reference encoders (CoordConv2d stack + BatchNorm2d + GRU, STL / cross attention) only run here
    "fs2_liu2021_train": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                              batch=2, s_max=20, s_step=6, pin=None, seed=12, prosody="liu2021"),
    # the other block types: dropout aside, training differs from eval in PostNet's (and conformer's) BatchNorm only
    "transformer_train": dict(dataset="LJSpeech", block_type="transformer", learn_alignment=False, mode="teacher",
                              batch=2, s_max=20, s_step=6, pin=None, seed=13),
    "fastformer_train": dict(dataset="LJSpeech", block_type="fastformer", learn_alignment=False, mode="teacher",
                             batch=2, s_max=20, s_step=6, pin=None, seed=14),
    "conformer_train": dict(dataset="LJSpeech", block_type="conformer", learn_alignment=False, mode="teacher",
                            batch=2, s_max=20, s_step=6, pin=None, seed=15),
    # unsupervised duration modelling in training (BASELINE configs[2] family: learn_alignment True): before
    # binarization_start_steps the decoder input is the SOFT alignment bmm(attn_soft, x) and gradients flow through the
    # aligner; afterwards MAS durations (no gradient through the hard path) drive the length regulator
    "fs2_unsup_train_soft": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=True, mode="unsup",
                                 batch=2, s_max=16, s_step=5, pin=None, seed=16, step=100),
    "conformer_unsup_train": dict(dataset="LJSpeech", block_type="conformer", learn_alignment=True, mode="unsup",
                                  batch=2, s_max=16, s_step=5, pin=None, seed=17, step=120000),
    # BASELINE configs[3] family: fastformer, VCTK (DeepSpeaker embeddings -> Linear, speaker-conditioned aligner)
    "fastformer_vctk_unsup_train": dict(dataset="VCTK", block_type="fastformer", learn_alignment=True, mode="unsup",
                                        batch=2, s_max=14, s_step=4, pin=None, seed=18, step=120000),
    "fs2_pitch_frame_train": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                                  batch=2, s_max=20, s_step=6, pin=None, seed=24, pitch_type="frame"),
    "fs2_pitch_ph_train": dict(dataset="LJSpeech", block_type="transformer_fs2", learn_alignment=False, mode="teacher",
                               batch=2, s_max=20, s_step=6, pin=None, seed=25, pitch_type="ph"),
}
CASES_ALL = dict(CASES, **TRAIN_CASES)
GRAD_SAMPLES = 512   # gradient entries stored per parameter tensor (evenly strided)

TAP_STRIDE = 4  # intermediate activations are stored for every 4th row only

# names of the arrays stored per case (prefix "ref.")
TUPLE_NAMES = ["mel", "postnet_mel", "p_predictions", "e_predictions", "log_d_predictions", "d_rounded", "src_masks",
               "mel_masks", "src_lens", "mel_lens", "attn_outs", "prosody_info", "p_targets", "e_targets"]


def build_case(name):
    """(configs, state_dict, batch) for a case -- importable without the reference."""
    from ctts_b200 import configs, spec, synth
    c = CASES_ALL[name]
    p, m, t = configs.builtin_configs(c["dataset"], block_type=c["block_type"], learn_alignment=c["learn_alignment"],
                                      prosody=c.get("prosody"), pitch_type=c.get("pitch_type"))
    entries, _, _ = spec.parameter_spec(p, m)
    sd = synth.synthetic_state_dict(entries, pin_frames_per_phoneme=c["pin"])
    batch = synth.ljspeech_batch(batch=c["batch"], s_max=c["s_max"], s_step=c["s_step"], mode=c["mode"],
                                 seed=c["seed"], spk_dim=512 if c["dataset"] == "VCTK" else None)
    return (p, m, t), sd, batch


def flatten_outputs(out):
    """14-tuple -> {name: numpy array} (None entries dropped, nested dict / tuple flattened with dots)."""
    import numpy as np
    import torch
    flat = {}

    def put(key, v):
        if v is None:
            return
        if isinstance(v, dict):
            for k in sorted(v):
                put(key + "." + k, v[k])
        elif isinstance(v, (tuple, list)):
            for i, x in enumerate(v):
                put("%s.%d" % (key, i), x)
        elif isinstance(v, torch.Tensor):
            flat[key] = v.detach().cpu().numpy()
        else:
            flat[key] = np.asarray(v)

    for n, v in zip(TUPLE_NAMES, out):
        put(n, v)
    return flat


def call_kwargs(batch, step=None):
    """Positional / keyword arguments of CompTransTTS.forward from a synth batch (deep-copied targets)."""
    import copy
    b = copy.deepcopy(batch)
    args = (b["speakers"], b["texts"], b["src_lens"], b["max_src_len"])
    kw = {k: b[k] for k in ("mels", "mel_lens", "max_mel_len", "p_targets", "e_targets", "d_targets", "attn_priors",
                            "spker_embeds") if k in b}
    if "attn_priors" in kw:
        kw["step"] = 120000 if step is None else step
    return args, kw


def train_objective(out):
    """A fixed scalar function of the differentiable outputs of the 14-tuple: every output is contracted with a
    deterministic pseudo-random tensor (cos of an index ramp), so all parameter gradients are exercised with O(1) weights."""
    import torch
    flat = []

    def put(v):
        if v is None:
            return
        if isinstance(v, dict):
            for k in sorted(v):
                put(v[k])
        elif isinstance(v, (tuple, list)):
            for x in v:
                put(x)
        elif isinstance(v, torch.Tensor) and v.is_floating_point() and v.requires_grad:
            flat.append(v)

    for i in (0, 1, 2, 3, 4, 10, 11):    # mel, postnet mel, pitch / energy / log-duration predictions, aligner
        # outputs (attn_soft, attn_logprob: what the reference's CTC / binarisation losses read), prosody info
        put(out[i])
    total = 0.0
    for j, v in enumerate(flat):
        w = torch.cos(torch.arange(v.numel(), dtype=torch.float32) * (0.37 + 0.11 * j)).reshape(v.shape)
        total = total + (v * w).sum() / (v.numel() ** 0.5)
    return total


def grad_sample_index(numel):
    """Indices (into the flattened gradient) stored in the fixtures."""
    import numpy as np
    if numel <= GRAD_SAMPLES:
        return np.arange(numel)
    return np.linspace(0, numel - 1, GRAD_SAMPLES).astype(np.int64)
