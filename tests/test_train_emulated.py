"""CPU: the HOST logic of the training step (ctts_b200/train_engine.py: tape, gradient routing, shapes, strides) with every
C-ABI call answered by the torch restatement in oracle/capi_emulator.py, against the reference's training fixtures
(tests/golden/*_train.npz: outputs, gradient samples + norms of every parameter, BatchNorm buffers).  The kernels
themselves are checked on the GPU (tests/test_gpu_train.py); this test is what can run without one."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
from oracle import capi_emulator as emu  # noqa: E402

import train_checks  # noqa: E402

BUILT = ["fs2_train", "fs2_unsup_train_soft", "transformer_train", "fastformer_train", "conformer_train",
         "conformer_unsup_train", "fastformer_vctk_unsup_train", "fs2_liu2021_train", "fs2_pitch_frame_train",
         "fs2_pitch_ph_train"]


@pytest.mark.parametrize("math_mode", ["tc", "fp32"])
@pytest.mark.parametrize("name", BUILT)
def test_training_step_host_logic(name, math_mode, golden_dir, monkeypatch):
    emu.install(monkeypatch)
    if math_mode == "fp32":
        monkeypatch.setenv("CTTS_DECODER_MATH", "fp32")
        monkeypatch.setenv("CTTS_ENCODER_MATH", "fp32")
        monkeypatch.setenv("CTTS_TRAIN_BWD_MATH", "fp32")
    monkeypatch.setenv("CTTS_DROPOUT", "0")
    net, out, loss = train_checks.run_training_case(name, torch.device("cpu"))
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    # fp32 arithmetic: the oracle's own tolerances (tests/test_oracle_train.py); tensor-core planes: north_star's output
    # tolerance (1e-3 abs / 1e-2 rel) and 3x the gradient allowance
    if math_mode == "fp32":
        train_checks.check_against_fixture(net, out, loss, gold)
    else:
        train_checks.check_against_fixture(net, out, loss, gold, out_atol=1e-3, out_rtol=1e-2, grad_scale=3.0, loss_rtol=1e-3,
                                           outlier_frac=0.03, outlier_bound=0.3)
    used = set(emu.CALLS)
    assert "ctts_layernorm_bwd" in used and "ctts_act_bwd" in used
    if math_mode == "tc":
        assert "ctts_gemm_wgrad" in used and "ctts_gemm_split" in used


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_inference_engine_host_logic(name, golden_dir, monkeypatch):
    """The INFERENCE engine (all block types; encoders on 3-plane, decoders on 2-plane tensor-core math, the conformer's
    padded-head batched plane GEMMs and plane-writing relative-shift softmax) with emulated kernels against the reference's
    eval-mode fixtures."""
    emu.install(monkeypatch)
    import ctts_b200
    (p, m, t), sd, batch = cases.build_case(name)
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd, strict=True)
    args, kw = cases.call_kwargs(batch)
    out = net(*args, **kw)
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    flat = cases.flatten_outputs(out)
    for k in ("mel", "postnet_mel"):
        a, b = gold["ref." + k], flat[k]
        assert a.shape == b.shape
        bad = np.abs(a - b) > 1e-3 + 1e-2 * np.abs(a)
        # fastformer is ill-conditioned by construction (its inverted mask rounds the logits to ulp(1e4)): flip budget, as
        # in tests/test_gpu_configs.py
        budget = 0.01 if "fastformer" in name else 0.0
        assert bad.mean() <= budget, "%s: %.3f%% outside tolerance, max err %.3g" % (k, 100 * bad.mean(), np.abs(a - b).max())
    assert np.array_equal(gold["ref.mel_lens"], flat["mel_lens"])
    if name.startswith("conformer"):
        assert "ctts_relshift_softmax_planes" in emu.CALLS and "ctts_gemm_batched_planes" in emu.CALLS


def test_engine_takes_the_fused_entry_points(monkeypatch):
    """fs2 inference: the decoder's projections carry their LayerNorm (ctts_gemm_split_ln), the encoder's attention is the
    fused short-sequence kernel, no V^T transpose is launched, and the weight gradient of the training step reads the
    row-major planes (ctts_gemm_wgrad_rowmajor) wherever Cin % 128 == 0."""
    emu.install(monkeypatch)
    import ctts_b200
    (p, m, t), sd, batch = cases.build_case("fs2_infer_c1")
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd, strict=True)
    args, kw = cases.call_kwargs(batch)
    del emu.CALLS[:]
    net(*args, **kw)
    used = list(emu.CALLS)
    n_layers = m["transformer_fs2"]["decoder_layer"]
    assert used.count("ctts_gemm_split_ln") == 2 * n_layers
    assert used.count("ctts_attention_small") == m["transformer_fs2"]["encoder_layer"]
    assert "ctts_transpose_v_planes" not in used and "ctts_attention_split" not in used
    monkeypatch.setenv("CTTS_DROPOUT", "0")
    del emu.CALLS[:]
    train_checks.run_training_case("fs2_train", torch.device("cpu"))
    used = set(emu.CALLS)
    assert "ctts_gemm_wgrad_rowmajor" in used and "ctts_gemm_wgrad" in used      # (the 80-channel mel convs keep the transposed path)


def test_parameter_table_follows_the_module(monkeypatch):
    """engine.Prepared keeps the name -> tensor table between calls (the tree walk costs 0.25 ms); it must notice every way
    the module can change under it: in-place updates (version counters), .to() / assigned state dicts (storages), and a
    Parameter object being replaced (structure epoch)."""
    emu.install(monkeypatch)
    import ctts_b200
    from ctts_b200.module import _Tracked
    (p, m, t), sd, batch = cases.build_case("fs2_infer_c1")
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd, strict=True)
    net.use_cuda_graphs = False          # (repeated calls: a second sighting of a shape would be captured into a CUDA graph)
    args, kw = cases.call_kwargs(batch)
    base = net(*args, **kw)[1].clone()
    prep = net._prepared
    sig = prep.sig
    assert torch.equal(net(*args, **kw)[1], base) and prep.sig is sig, "an unchanged module must not be re-laid-out"
    # in-place update (an optimizer step)
    with torch.no_grad():
        net.get_parameter("mel_linear.bias").add_(1.0)
    out = net(*args, **kw)[1]
    assert prep.sig is not sig and not torch.allclose(out, base)
    # a Parameter object replaced: the cached table would still point at the old tensor
    sig, epoch = prep.sig, _Tracked.structure_epoch
    node = net.get_submodule("mel_linear")
    node.bias = torch.nn.Parameter(node.bias.detach() - 1.0)
    assert _Tracked.structure_epoch > epoch
    out = net(*args, **kw)[1]
    assert prep.sig is not sig
    np.testing.assert_allclose(out.numpy(), base.numpy(), atol=1e-5, rtol=1e-5)
    # load_state_dict(assign=True) swaps the Parameter objects as well
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["mel_linear.bias"] = sd2["mel_linear.bias"] + 2.0
    net.load_state_dict(sd2, strict=True, assign=True)
    out = net(*args, **kw)[1]
    assert not torch.allclose(out, base)
